#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 900 > gpurun_out/tests_multi_n2.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/tests_multi_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_n2.err
python - <<'PY'
import json
b = json.load(open("gpurun_out/bench_n2.json"))
print("c3 N=2 value %.4g ms/step %.3f e2e %.4g (%.3f ms) frac %.3f golden %s" % (b["value"], b["ms_per_step"], b["e2e"]["value"], b["e2e"]["ms_per_step"], b["roofline"]["frac"], b["config"]["golden_check"]))
for k in ("kriging", "kriging_c5"):
    d = b[k]; print(k, "%.4g" % d["value"], d["ms_per_step"], d["phases_ms"], "e2e %.4g" % d["e2e"]["value"])
PY
