"""CPU tests of the drop-in boundary: libarmsim.so builds for sm_100a, loads, exports every symbol include/armsim.h
declares, the ctypes mirrors match the C structs, and the product path fails LOUDLY without a GPU (no fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "armsim.h")).read()
    declared = sorted(set(re.findall(r"\b(armsim_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 17
    lib = pkg._lib.lib()
    for sym in declared:
        assert hasattr(lib, sym), "libarmsim.so does not export %s" % sym
    assert sorted(pkg._lib.EXPORTS) == declared
    assert lib.armsim_abi_version() == pkg._lib.ABI_VERSION == 5


def test_library_is_sm100a_only(pkg):
    out = subprocess.run(["cuobjdump", "--list-elf", pkg._build.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_config_struct_matches_c_layout(pkg, oracle, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "armsim.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(ArmsimConfig),offsetof(ArmsimConfig,dv),offsetof(ArmsimConfig,init_q),'
                   'offsetof(ArmsimConfig,custom_chain),sizeof(ArmsimChain));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    size, off_dv, off_q, off_chain, chain = map(int, subprocess.check_output([str(exe)]).split())
    for mod in (pkg._lib, oracle):
        assert C.sizeof(mod.ArmsimConfig) == size
        assert mod.ArmsimConfig.dv.offset == off_dv
        assert mod.ArmsimConfig.init_q.offset == off_q
        assert mod.ArmsimConfig.custom_chain.offset == off_chain
        assert C.sizeof(mod.ArmsimChain) == chain


@pytest.mark.parametrize("task,dv,dis,steps,zhi", [
    (0, 0.02, 0.01, 500, 0.55),        # config.py:41,42,51; rl_reach_env.py:223
    (1, 0.08, 0.05, 500, 0.1),         # rl_push_env.py:322,86,314
    (2, 0.08, 0.05, 500, 0.807),       # rl_pick_env.py:321,313
    (3, 0.005, 0.1, 1000, 0.55),       # kuka_reach_env.py:215,289,59
])
def test_default_config_is_the_reference_constants(pkg, oracle, task, dv, dis, steps, zhi):
    a = pkg._lib.default_config(task)
    b = oracle.default_config(task)
    assert bytes(a)[:C.sizeof(a) - 8] == bytes(b)[:C.sizeof(b) - 8]          # product and oracle agree field for field
    assert a.dv == dv and a.reach_dis == dis and a.max_steps == steps and a.ws_hi[2] == pytest.approx(zhi)
    assert list(a.ws_lo) == [0.2, -0.3, 0.0] and list(a.ws_hi)[:2] == [0.7, 0.3]
    assert list(a.init_q) == [0.006418, 0.413184, -0.011401, -1.589317, 0.005379, 1.137684, -0.006539]
    assert a.ik_damping == 1e-5 and a.ik_max_iters == 20 and a.ik_residual == 1e-4
    assert a.struct_size == C.sizeof(a)


def test_bad_arguments_return_error_codes(pkg):
    L = pkg._lib
    lib = L.lib()
    cfg = L.default_config(0)
    h = C.c_void_p()
    cfg.struct_size = 8
    assert lib.armsim_create(C.byref(cfg), C.byref(h)) == -1 and b"struct_size" in lib.armsim_last_error()
    cfg = L.default_config(0, n_envs=0)
    assert lib.armsim_create(C.byref(cfg), C.byref(h)) == -1
    cfg = L.default_config(0, mapping=2)      # the four-lanes-per-arm mapping measured in round 2 was removed (armsim.h)
    assert lib.armsim_create(C.byref(cfg), C.byref(h)) == -1 and b"mapping" in lib.armsim_last_error()
    assert lib.armsim_default_config(9, C.byref(cfg)) == -1
    assert lib.armsim_step(None, None, None, None, None, None, None) == -1
    assert lib.armsim_obs_dim(None) == -1


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the product path must raise, not silently compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.ArmsimError, match="no usable CUDA device|no CPU fallback"):
        pkg.ArmSimHandle("reach", n_envs=4)
    with pytest.raises(pkg.ArmsimError):
        pkg.RLReachEnv()
    with pytest.raises(pkg.ArmsimError):
        pkg.BatchedArmEnv("reach", n_envs=4)


def test_product_never_imports_oracle():
    """rule: nothing under drl-on-robot-arm_b200/ may import, link or call oracle/."""
    pkg_dir = os.path.join(ROOT, "drl-on-robot-arm_b200")
    for dp, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "liboracle" not in txt and "armsim_oracle" not in txt, f
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f
    out = subprocess.run(["ldd", os.path.join(pkg_dir, "libarmsim.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_env_registry_names(pkg):
    """envs/__init__.py:1-3 + kuka_reach_env.py: lookup by name as main.py:83 does"""
    for name in ("RLReachEnv", "RLPushEnv", "RLPickEnv", "KukaReachEnv"):
        cls = getattr(pkg.envs, name)
        assert callable(cls)
    assert pkg.opt.reach_ctr == 0.02 and pkg.opt.reach_dis == 0.01 and pkg.opt.max_steps_one_episode == 500
