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


def source_hash():
    """sha256 over the library's sources (csrc/* and include/smct.h, by file name and content).  The hash is compiled
    into the .so (smct_source_hash()) so that a binary can be matched to the tree it was built from."""
    import hashlib
    h = hashlib.sha256()
    for d in sorted(_deps()):
        h.update(os.path.basename(d).encode() + b"\0")
        with open(d, "rb") as f:
            h.update(f.read())
        h.update(b"\0")
    return h.hexdigest()[:32]


def built_hash():
    """Source hash embedded in the existing .so (read from the binary's bytes: no dlopen), or None."""
    if not os.path.exists(LIB):
        return None
    with open(LIB, "rb") as f:
        blob = f.read()
    tag = b"smct-source-hash:"
    i = blob.find(tag)
    if i < 0:
        return None
    return blob[i + len(tag):i + len(tag) + 32].decode("ascii", "replace")


def needs_build():
    """True when there is no library or it was compiled from different sources (hash, not mtime)."""
    return built_hash() != source_hash()


def build(force=False, verbose=False):
    """Compile the library for sm_100a.  Returns the path of the .so.  Several processes of one job (torchrun ranks) may
    call this at once: the compile runs under an exclusive file lock and the others re-check the hash afterwards."""
    if not force and not needs_build():
        return LIB
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():   # another process built it while this one waited
                return LIB
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("libsmct_b200.so is missing or was built from other sources (embedded hash %s, tree %s) and nvcc "
                           "is not available to rebuild it" % (built_hash(), source_hash()))
    tmp = LIB + ".tmp.%d" % os.getpid()
    cmd = [nvcc] + NVCC_FLAGS + ['-DSMCT_SOURCE_HASH="%s"' % source_hash(), "-o", tmp] + SOURCES
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode == 0:
        os.replace(tmp, LIB)       # atomic: a concurrently loading process never sees a half-written library
    elif os.path.exists(tmp):
        os.remove(tmp)
    cmd[-len(SOURCES) - 1] = LIB
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
