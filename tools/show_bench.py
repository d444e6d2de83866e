"""Print the headline numbers of a bench.py JSON line."""
import json, sys
b = json.load(open(sys.argv[1]))
def show(d, name):
    cb = d.get("cpu_baseline", {})
    print(f"{name:11s} value {d['value']:.4g} {d['unit']}  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:.4g} "
          f"({d['e2e']['ms_per_step']:.3f} ms)  frac {d['roofline']['frac']:.3f}  kernel_ms {d['roofline']['kernel_ms']:.3f} "
          f"{d.get('phases_ms', '')}  cpu {cb.get('value')} {cb.get('kind')} x{cb.get('cores')}")
show(b, "c3 N=%d" % b["n_gpus"])
for k in ("c2", "kriging", "kriging_c5"):
    if k in b:
        show(b[k], k)
print("golden:", b["config"]["golden_check"], " clocks:", b["clocks"])
