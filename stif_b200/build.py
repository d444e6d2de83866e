"""Build libstif_b200.so (sm_100a CUDA, C ABI) in-tree with nvcc.

`python -m stif_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without
a GPU.  The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libstif_b200.so")
SOURCES = ["context.cu", "variogram.cu", "kriging.cu"]
HEADERS = ["common.cuh", "krig_chol.cuh", "krig_ldl.cuh", os.path.join("..", "..", "include", "stif_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--fmad=false",          # every FMA on the parity path is written explicitly
    "-cudart", "static",
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libstif_b200.so")
    return nvcc


def _digest():
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        p = os.path.join(CSRC, name)
        if os.path.exists(p):
            with open(p, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_variant(out_path, extra_flags):
    """Development aid: build another copy of the library with extra nvcc flags (e.g.
    -DSTIF_VARIO_TJ=256) for A/B timing through tools/*.py's STIF_LIB switch."""
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build", "variant_" + hashlib.sha256(" ".join(extra_flags).encode()).hexdigest()[:8])
    os.makedirs(objdir, exist_ok=True)
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src}")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static",
                           "-o", out_path] + objs)
    return out_path


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library. Returns its path."""
    stamp = LIB_PATH + ".stamp"
    digest = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return LIB_PATH
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", sp, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static",
           "-o", LIB_PATH] + objs
    subprocess.check_call(cmd)
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    if "--variant" in sys.argv:      # python -m stif_b200.build --variant out.so -DSTIF_VARIO_TJ=256 ...
        i = sys.argv.index("--variant")
        print(build_variant(os.path.abspath(sys.argv[i + 1]), sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
