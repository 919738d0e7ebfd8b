"""Build libedge_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the repo snapshot.  Flags that are part of
the numerical contract (DESIGN.md "Canonical arithmetic"): -fmad=false (FMAs only where fmaf() is
written), IEEE division / sqrt, no flush-to-zero, no fast-math.
"""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libedge_b200.so")
SOURCES = ["ee_capi.cu"]
HEADERS = ["ee_device.cuh", "ee_edge_step125.cuh", "ee_edge_canny.cuh", "ee_edge_fast.cuh", "ee_edge_canny_fast.cuh", "ee_attack.cuh", "ee_square.cuh", "ee_hfs.cuh", "ee_edge_cluster.cuh", "ee_edge_tiles.cuh", "ee_edge_canny_tiles.cuh", "ee_edge_stream.cuh", "ee_gf.cuh", "ee_pgd_l2.cuh", "ee_hfs_tc.cuh",
           os.path.join("..", "..", "include", "edge_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-diag-suppress", "177",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libedge_b200.so cannot be built")
    return exe


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


PARTS = (1, 2, 3, 4, 5, 6)      # ee_capi.cu is compiled once per kernel family (-DEE_PART=k), in parallel, then linked


def build(force=False, verbose=False, extra_flags=()):
    """Compile if missing or older than its sources.  Returns the path of the .so.

    Safe under torchrun / DataParallel on a fresh checkout: an exclusive file lock serialises the builders, each one
    compiles into its own temporary directory and the finished library is moved onto its final name with os.replace(),
    so no process ever dlopen()s a half-written file; whoever gets the lock second finds the library fresh and returns."""
    if not force and not is_stale():
        return LIB
    import fcntl
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    nvcc = _nvcc()
    os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
    with open(os.path.join(PKG, "build", ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():          # another process built it while we waited
                return LIB
            objdir = tempfile.mkdtemp(prefix="obj_%d_" % os.getpid(), dir=os.path.join(PKG, "build"))
            compile_flags = [f for f in NVCC_FLAGS if f != "-shared"] + list(extra_flags)
            src = os.path.join(CSRC, SOURCES[0])

            def compile_part(k):
                obj = os.path.join(objdir, "ee_part%d.o" % k)
                cmd = [nvcc] + compile_flags + ["-DEE_PART=%d" % k, "-c", src, "-o", obj]
                if verbose:
                    print(" ".join(cmd), flush=True)
                subprocess.check_call(cmd)
                return obj

            try:
                with ThreadPoolExecutor(max_workers=min(len(PARTS), os.cpu_count() or 1)) as pool:
                    objs = list(pool.map(compile_part, PARTS))
                tmp_lib = os.path.join(objdir, "libedge_b200.so")
                link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"] + objs + ["-o", tmp_lib]
                if verbose:
                    print(" ".join(link), flush=True)
                subprocess.check_call(link)
                os.replace(tmp_lib, LIB)
            finally:
                shutil.rmtree(objdir, ignore_errors=True)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose=True,
                extra_flags=["-Xptxas", "-v"] if "-v" in sys.argv else ()))
