"""Top stalled SASS instructions of a kernel from an .ncu-rep (source page).
usage: python tools/ncu_hot.py rep kernel_regex [topN]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
blocks = raw.split('"Kernel Name",')
blk = blocks[1]
lines = blk.split("\n")
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[1:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]]) for r in data)
totinst = sum(int(r[ix["Instructions Executed"]]) for r in data)
print("kernel", lines[0][:80], "total samples", tot, "total warp-instr", totinst)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
cum = 0
ranked = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:top]
for i in sorted(ranked):
    r = data[i]
    s = int(r[ix["# Samples"]])
    main = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"{i:5d} {100*s/tot:5.1f}% inst={int(r[ix['Instructions Executed']]):>10d} thr={r[ix['Avg. Threads Executed']]:>5s} "
          f"{main[0][1]}:{main[0][0]} {main[1][1]}:{main[1][0]}  | {r[ix['Source']].strip()[:70]}")
