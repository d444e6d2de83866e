"""Executed warp-instructions and stall samples per CUDA source line (needs -lineinfo and
--import-source on).  usage: python tools/ncu_lines.py rep kernel_regex [topN]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda",
                      "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
agg = {}
hdr = None
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        key = (fname, int(r[0]), r[1].strip()[:90])
        try:
            s, i = int(r[i_s]), int(r[i_i])
        except ValueError:
            continue
        a = agg.setdefault(key, [0, 0])
        a[0] += s
        a[1] += i
ts = sum(v[0] for v in agg.values()) or 1
ti = sum(v[1] for v in agg.values()) or 1
print(f"total samples {ts}, warp-instr {ti}")
for (f, ln, src), (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*i/ti:5.1f}% instr {100*s/ts:5.1f}% samples  {f}:{ln:<4d} {src}")
