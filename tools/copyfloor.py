"""ctypes loader / builder of tools/copyfloor.cu (the pure-CUDA host<->device copy floor used by bench.py's e2e leg).
A measuring tool: it allocates its own buffers and is not part of libedge_b200.so."""
import ctypes
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "copyfloor.cu")
LIB = os.path.join(HERE, "libcopyfloor.so")


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    tmp = LIB + ".%d.tmp" % os.getpid()
    subprocess.check_call([nvcc, "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared",
                           SRC, "-o", tmp])
    os.replace(tmp, LIB)
    return LIB


_L = None


def measure(device, nbytes, chunks=1, iters=10, mode=2, host_register=False):
    """ms per iteration of `nbytes` H2D (mode 0), D2H (1) or both concurrently (2) on `device`."""
    global _L
    if _L is None:
        _L = ctypes.CDLL(build())
        _L.cf_measure.argtypes = [ctypes.c_int, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                  ctypes.POINTER(ctypes.c_double)]
    ms = ctypes.c_double(0.0)
    rc = _L.cf_measure(int(device), int(nbytes), int(chunks), int(iters), int(mode), int(bool(host_register)), ctypes.byref(ms))
    if rc != 0:
        raise RuntimeError("copyfloor: CUDA error %d" % rc)
    return ms.value
