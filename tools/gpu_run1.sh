#!/bin/bash
# GPU run: full -m gpu suite (no -x, to see every failure), smoke, a short bench line.
mkdir -p gpurun_out
rm -f gpurun_out/parity_quantiles.jsonl
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/tests.log
tail -30 gpurun_out/tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    b = json.load(open("gpurun_out/bench.json"))
    def show(d, name):
        print(name, "value %.4g %s" % (d["value"], d["unit"]), "ms/step %.3f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"],
              "frac %.3f" % d["roofline"]["frac"], "kernel_ms %.3f" % d["roofline"]["kernel_ms"], d.get("phases_ms"), d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("kind"))
    show(b, "c3")
    for k in ("c2", "kriging", "kriging_c5"):
        if k in b: show(b[k], k)
except Exception as e:
    print("bench parse failed", e)
PY
