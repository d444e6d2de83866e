#!/bin/bash
# Evidence pass on the current build: profiles, full GPU suite + bench (N = 1), then N = 2, 4, 8 bench lines.
set -x
cd "$(dirname "$0")/.."
/usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_profile.sh' > /tmp/fp_profile.log 2>&1
tail -3 /tmp/fp_profile.log
bash tools/collect_profiles.sh
/usr/local/graft/bin/gpurun --timeout 3000 -- 'bash tools/gpu_run1.sh' > /tmp/fp_run1.log 2>&1
tail -12 /tmp/fp_run1.log
SUF=${SUF:-g}
cp gpurun_out/tests.log profiles/r02_tests_$SUF.log
cp gpurun_out/bench.json profiles/r02_bench_n1_$SUF.json
cp gpurun_out/parity_quantiles.jsonl profiles/r02_parity_quantiles.jsonl
for N in 2 4 8; do
  /usr/local/graft/bin/gpurun --gpus $N --timeout 900 -- "N=$N bash tools/gpu_run_n8.sh" 2>&1 | grep "c3 N\|kriging\|GPU-minutes\|status"
  cp gpurun_out/bench_n$N.json profiles/r02_bench_n${N}_c.json
done
