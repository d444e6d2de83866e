"""Opcode histogram of the hot kernels' SASS (cuobjdump), the evidence for the instruction
selection DESIGN.md describes: packed FP32 (FADD2/FMUL2/FFMA2), bulk copies (UBLKCP) and
mbarriers (SYNCS), FP64 tensor-core DMMA, native L2 / shared atomics (REDG / ATOMS).
usage: python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "stif_b200", "libstif_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KERNELS = ["pair_kernelILi0ELb0", "pair_kernelILi0ELb1", "solve_chol_kernelILi2ELi2ELi0", "search_kernelILi1", "solve_blk_kernelILi4ELi2", "solve_ldl_kernelILi4ELi4ELi2", "solve_ldl_kernelILi2ELi1ELi2",
           "solve_minnorm_kernel", "reduce_kernelILb1"]
INTEREST = ("FADD2", "FMUL2", "FFMA2", "UBLKCP", "SYNCS", "DMMA", "REDG", "ATOMS", "ATOMG", "DFMA", "DADD", "DMUL", "DSETP", "MUFU",
            "LDS", "STS", "LDG", "STG", "SHFL", "VOTE", "REDUX", "BAR", "F2I", "I2F")
cur, hist = None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
print("library:", os.path.relpath(lib, ROOT), " build", open(lib + ".stamp").read().strip()[:16])
for key in KERNELS:
    for name, h in hist.items():
        if key in name:
            base = collections.Counter()
            for op, n in h.items():
                base[op.split(".")[0]] += n
            print("=" * 100)
            print(subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:160], f"  ({sum(h.values())} instructions)")
            print("  by mnemonic:", ", ".join(f"{op} {n}" for op, n in base.most_common(28)))
            sel = {op: n for op, n in h.items() if op.split(".")[0] in INTEREST and op.split(".")[0] in
                   ("FADD2", "FMUL2", "FFMA2", "UBLKCP", "SYNCS", "DMMA", "REDG", "ATOMS", "ATOMG", "MUFU", "REDUX")}
            print("  of interest:", ", ".join(f"{op} x{n}" for op, n in sorted(sel.items())))
