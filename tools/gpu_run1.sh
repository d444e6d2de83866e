#!/bin/bash
# GPU run: full -m gpu suite, smoke, a bench line.
mkdir -p gpurun_out
rm -f gpurun_out/parity_quantiles.jsonl
python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/tests.log
tail -8 gpurun_out/tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err
python tools/show_bench.py gpurun_out/bench.json
