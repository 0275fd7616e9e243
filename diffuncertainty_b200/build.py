"""Build libvalunc.so in-tree with nvcc for sm_100a (no JIT cache, no pip).

    python -m diffuncertainty_b200.build [--force] [--verbose]

The shared library lands in diffuncertainty_b200/lib/ so it travels with the
repository snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libvalunc.so")
SOURCES = ["abi.cu", "k1_fused.cu", "k1_tma.cu", "k1_uni.cu", "k1_co_tma.cu", "k2_spatial.cu", "k3_map_stats.cu", "k4_quantile.cu", "k5_members.cu", "synth.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _digest() -> str:
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + ["../../include/valunc.h"]
    for n in names:
        p = os.path.join(CSRC, n)
        if os.path.isfile(p):
            h.update(n.encode())
            with open(p, "rb") as f:
                h.update(f.read())
    # the flags without the checkout's absolute paths: the library built in one place (the build container) must be
    # recognised as current where the tree is copied to (the GPU box) instead of being rebuilt by every process there
    h.update(" ".join(f.replace(ROOT, "<root>") for f in FLAGS).encode())
    return h.hexdigest()


def _current(stamp: str, digest: str) -> bool:
    try:
        return os.path.isfile(LIB) and open(stamp).read().strip() == digest
    except OSError:
        return False


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile and link libvalunc.so if it is missing or older than its sources.  Safe to call from many processes at once
    (the ranks of a torchrun job): one of them builds under a file lock, into a temporary file that is renamed into place."""
    import fcntl
    os.makedirs(OBJDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.stamp")
    digest = _digest()
    if not force and _current(stamp, digest):
        return LIB
    with open(os.path.join(LIBDIR, "build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _current(stamp, digest):  # another process built it while this one waited
                return LIB
            return _build_locked(stamp, digest, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(stamp: str, digest: str, verbose: bool) -> str:
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose:
            sys.stderr.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    # the CUDA runtime is linked SHARED: inside a PyTorch process the library must use the same runtime instance as
    # torch (one notion of the current device and of the streams torch hands over); a private static runtime
    # launches on device 0 whatever torch.cuda.current_device() is.  libcudart.so.12 is already in the process
    # when torch is imported; the rpath covers stand-alone C users.
    tmp = LIB + f".tmp{os.getpid()}"
    cmd = [NVCC, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "shared",
           "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr}")
    os.replace(tmp, LIB)  # atomic: a process that loads the library concurrently sees the old or the new file, never half of one
    with open(stamp + ".tmp", "w") as f:
        f.write(digest)
    os.replace(stamp + ".tmp", stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
