"""Aggregate executed warp-instructions and stall samples over SASS index regions."""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
step = int(sys.argv[3]) if len(sys.argv) > 3 else 100
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
blk = raw.split('"Kernel Name",')[1]
lines = blk.split("\n")
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[1:] if len(r) == len(hdr)]
tot_s = sum(int(r[ix["# Samples"]]) for r in data); tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data)
print("instrs", len(data), "samples", tot_s, "warp-instr", tot_i)
for b in range(0, len(data), step):
    seg = data[b:b + step]
    s = sum(int(r[ix["# Samples"]]) for r in seg); i = sum(int(r[ix["Instructions Executed"]]) for r in seg)
    if s / tot_s > 0.01 or i / tot_i > 0.01:
        ops = {}
        for r in seg:
            op = r[ix["Source"]].split()[0] if not r[ix["Source"]].strip().startswith("@") else r[ix["Source"]].split()[1]
            ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + int(r[ix["Instructions Executed"]])
        topops = ", ".join(f"{k}:{v*100//max(i,1)}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:5])
        print(f"[{b:5d},{b+step:5d}) samples {100*s/tot_s:5.1f}%  instr {100*i/tot_i:5.1f}%   {topops}")
