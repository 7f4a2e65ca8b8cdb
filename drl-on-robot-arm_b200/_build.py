"""Build recipe for libarmsim.so (hand-written CUDA for sm_100a behind the C-ABI of include/armsim.h).

nvcc cross-compiles without a GPU; the .so is built IN-TREE (git-ignored, but it travels to the GPU box).
"""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_PATH = os.path.join(PKG, "libarmsim.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only: no multi-arch fatbin, no PTX fallback
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "--cudart", "static",                           # the .so loads on a box without libcudart on the path
]


def _sources():
    src = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    src += [os.path.join(ROOT, "include", f) for f in sorted(os.listdir(os.path.join(ROOT, "include")))]
    return src


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources())


def build_libarmsim(force=False, verbose=False):
    """Compile csrc/*.cu into libarmsim.so.  Raises if nvcc is missing or the compile fails."""
    if not force and not needs_build():
        return LIB_PATH
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:      # one builder at a time (torchrun ranks share the tree)
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not needs_build():          # another process built it while we waited
            return LIB_PATH
        return _build_locked(verbose)


def _build_locked(verbose):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libarmsim.so (no CPU fallback exists)")
    cus = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    cmd = [nvcc] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", LIB_PATH + ".tmp"] + cus
    if os.environ.get("ARMSIM_LANE_BLOCK"):          # tuning experiments only (reach: cube scratch assumes 128)
        cmd.insert(1, "-DARMSIM_LANE_BLOCK=" + os.environ["ARMSIM_LANE_BLOCK"])
    if os.environ.get("ARMSIM_SPARSE_MIN_BLOCKS"):   # tuning experiments only
        cmd.insert(1, "-DARMSIM_SPARSE_MIN_BLOCKS=" + os.environ["ARMSIM_SPARSE_MIN_BLOCKS"])
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_libarmsim(force=True, verbose=True))
