"""In-tree build of libsmct_b200.so (hand-written sm_100a CUDA kernels + the C-ABI of include/smct.h).

`python smc-t-v2_b200/build.py` or `build()`; nvcc cross-compiles without a GPU.  The .so stays in-tree
(git-ignored) so that it travels to the GPU box with the repository snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsmct_b200.so")
SOURCES = [os.path.join(CSRC, "smct_capi.cu")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def _deps():
    deps = [os.path.join(ROOT, "include", "smct.h")]
    for f in os.listdir(CSRC):
        if f.endswith((".cu", ".cuh", ".h")):
            deps.append(os.path.join(CSRC, f))
    return deps


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=False):
    """Compile the library for sm_100a.  Returns the path of the .so."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        if os.path.exists(LIB):
            return LIB  # GPU box without a toolchain: use the prebuilt library that travelled with the snapshot
        raise RuntimeError("nvcc not found and libsmct_b200.so is not built")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + SOURCES
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed (exit %d); see %s" % (r.returncode, log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
