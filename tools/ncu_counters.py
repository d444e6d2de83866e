"""Extract per-launch counters of named kernels from .ncu-rep captures into
profiles/r02_kernel_counters.json (read by bench.py for `roofline.traffic` / `roofline.executed`).

usage: python tools/ncu_counters.py key=report.ncu-rep[:kernel_regex] ...
Each entry records the digest of the CUDA sources the library was built from (the build stamp),
so a reader can see whether the capture matches the library that ran."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r02_kernel_counters.json")
METRICS = {"warp_instr": "smsp__inst_executed.sum", "lsu_wavefronts": "l1tex__data_pipe_lsu_wavefronts.sum",
           "shared_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
           "duration_ns": "gpu__time_duration.sum", "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "fp64_pipe_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
           "lsu_pipe_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "l1tex_pct": "l1tex__throughput.avg.pct_of_peak_sustained_active",
           "registers": "launch__registers_per_thread", "smem_dynamic": "launch__shared_mem_per_block_dynamic"}
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e6, "us": 1e3, "ns": 1.0, "s": 1e9}
try:
    with open(os.path.join(ROOT, "stif_b200", "libstif_b200.so.stamp")) as f:
        build = f.read().strip()[:16]
except OSError:
    build = None
try:
    out = json.load(open(OUT))
except Exception:
    out = {}
for arg in sys.argv[1:]:
    key, spec = arg.split("=", 1)
    rep, _, pat = spec.partition(":")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    sel = [r for r in rows[2:] if (not pat) or pat in r[ix["Kernel Name"]]]
    if not sel:
        print("no kernel matching", pat, "in", rep)
        continue
    r = sel[0]
    ent = {"kernel": r[ix["Kernel Name"]], "source": "profiles/" + os.path.basename(rep).replace(".ncu-rep", "_summary.txt"),
           "build": build}
    for name, metric in METRICS.items():
        if metric in ix:
            val = float(r[ix[metric]].replace(",", ""))
            ent[name] = val * UNIT.get(units[ix[metric]], 1.0) if name.startswith(("dram", "duration", "smem")) else val
    ent["dram_bytes"] = ent.get("dram_read", 0.0) + ent.get("dram_write", 0.0)
    out[key] = ent
    print(key, json.dumps(ent))
json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
