"""Shared-memory wavefronts (actual / ideal) and executed warp-instructions per CUDA source
line.  usage: python tools/ncu_wavefronts.py rep kernel_regex [topN]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda",
                      "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
agg, hdr, fname = {}, None, ""
for r in csv.reader(io.StringIO(raw)):
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        cols = [hdr.index(c) for c in ("Instructions Executed", "L1 Wavefronts Shared",
                                       "L1 Wavefronts Shared Ideal", "# Samples")]
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        key = (fname, int(r[0]), r[1].strip()[:80])
        try:
            vals = [int(r[c] or 0) for c in cols]
        except ValueError:
            continue
        a = agg.setdefault(key, [0, 0, 0, 0])
        for k in range(4):
            a[k] += vals[k]
tw = sum(v[1] for v in agg.values()) or 1
ti = sum(v[0] for v in agg.values()) or 1
print(f"warp-instr {ti}, shared wavefronts {tw}, ideal {sum(v[2] for v in agg.values())}")
for (f, ln, src), (i, w, wi, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*w/tw:5.1f}% wavefronts ({w:>10d}, ideal {wi:>10d}) {100*i/ti:5.1f}% instr  {f}:{ln:<4d} {src}")
