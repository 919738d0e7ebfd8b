"""Compile oracle/ee_oracle.c into oracle/_build/libee_oracle_{fma,generic}.so (gcc only).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Two builds of the same source:
  * ``fma``     : -mavx2 -mfma, fmaf() inlines to one vfmadd (fast; needs an FMA-capable host)
  * ``generic`` : baseline x86-64, fmaf() goes through libm (still correctly rounded)
Both use -ffp-contract=off so that fused multiply-adds happen only where fmaf() is written;
the two builds are therefore bit-identical and the loader picks whichever the host supports.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
SRC = os.path.join(HERE, "ee_oracle.c")

COMMON = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
          "-fno-fast-math", "-fno-unsafe-math-optimizations", "-Wall", "-Wextra", "-Wno-unused-parameter"]


def lib_path(kind):
    return os.path.join(OUT, "libee_oracle_%s.so" % kind)


def build(force=False, verbose=False):
    os.makedirs(OUT, exist_ok=True)
    built = []
    for kind, extra in (("fma", ["-mavx2", "-mfma"]), ("generic", [])):
        dst = lib_path(kind)
        if (not force and os.path.exists(dst)
                and os.path.getmtime(dst) >= os.path.getmtime(SRC)):
            built.append(dst)
            continue
        cmd = COMMON + extra + [SRC, "-o", dst, "-lm"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        built.append(dst)
    return built


if __name__ == "__main__":
    for p in build(force="--force" in sys.argv, verbose=True):
        print("built", p)
