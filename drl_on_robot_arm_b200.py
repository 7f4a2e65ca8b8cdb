"""Import alias: `import drl_on_robot_arm_b200` loads the package that lives in ./drl-on-robot-arm_b200/
(the directory keeps the project's hyphenated name, which is not a Python identifier)."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "drl-on-robot-arm_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_pkg_dir, "__init__.py"),
                                               submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
