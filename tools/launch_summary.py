"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): total time, share and
count per kernel.  usage: python tools/launch_summary.py launches.csv [header line]"""
import csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}.get(r[iu], 1e-6)
    name = re.sub(r"\(.*", "", r[ik]).replace("stif::", "")
    a = agg.setdefault(name, [0.0, 0])
    a[0] += v * scale
    a[1] += 1
tot = sum(a[0] for a in agg.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
print("(cold-cache, serialised launches: shares, not absolute times)")
for name, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:16]:
    print(f"{ms:9.3f} ms {100*ms/tot:5.1f}%  x{n:3d}  ({ms/n:.3f} ms each)  {name[:90]}")
