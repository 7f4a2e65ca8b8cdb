"""The device cube / contact model compiled FOR THE HOST (same source: csrc/cube_model.cuh, fp32, look-ahead
Gauss-Seidel) against the oracle's plain fp64 restatement (oracle/cube_model.h), without a GPU: g++ builds
tests/host/cube_host_check.cpp, which steps both from identical f32-rounded states (cubes resting, tumbling, pushed
sideways, squeezed between a capsule and the table, gripped) and fails unless >= 99.5 % of the sim steps agree to
5e-5 m / 5e-3 m/s (the tolerance of tests/test_parity_gpu.py::test_cube_tasks_teacher_forced; a contact entering the
5 mm margin on one side only is a legitimate one-step difference)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_device_cube_source_matches_oracle_on_host(tmp_path):
    exe = str(tmp_path / "cube_host_check")
    subprocess.check_call(["g++", "-O2", "-I", os.path.join(ROOT, "drl-on-robot-arm_b200", "csrc"), "-I", os.path.join(ROOT, "oracle"),
                           os.path.join(ROOT, "tests", "host", "cube_host_check.cpp"), "-o", exe, "-lm"])
    res = subprocess.run([exe, "300"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    lines = res.stdout.strip().splitlines()
    assert len(lines) == 2 and lines[0].startswith("push") and lines[1].startswith("pick")
    for ln in lines:
        f = ln.split()
        stats = dict(zip(f[1::2], f[2::2]))
        assert int(stats["squeezed"]) > 100 and int(stats["touching"]) > 100, ln    # the hard cases were really exercised
