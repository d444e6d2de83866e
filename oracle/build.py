"""Compile the C restatement (oracle/*.c) into oracle/_build/liboracle.so with gcc.

TEST INFRASTRUCTURE.  -ffp-contract=off: every fused multiply-add the reference's
LLVM code executes is written as an explicit fma() in the sources.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB_PATH = os.path.join(OUT_DIR, "liboracle.so")
SOURCES = ["oracle_variogram.c", "oracle_kriging.c"]
CFLAGS = ["-O2", "-mfma", "-mavx2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp",
          "-shared", "-fPIC"]


def build(force=False):
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(s) for s in srcs)
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.check_call(["gcc"] + CFLAGS + ["-o", LIB_PATH] + srcs + ["-lm", "-ldl"])
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True))
