#!/bin/bash
# Round-2 evidence run: launch list of the bench, one ncu --set full capture per hot kernel,
# compute-sanitizer memcheck + racecheck.  Outputs under gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
cap() { # name kernel_regex skip env... -- cmd
  name=$1; regex=$2; skip=$3; shift 3
  env "$@" ncu --set full --import-source on --clock-control none -k "regex:$regex" -s $skip -c 1 -f -o gpurun_out/r02_ncu_$name \
      python $PROG > gpurun_out/ncu_$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/ncu_$name.log)"
}
PROG=tools/vario_time.py cap pair_c3 pair_kernel 1 SIZES=1000000 REPS=2
PROG=tools/vario_time.py cap pair_c2 pair_kernel 1 SIZES=100000 REPS=2
PROG=tools/krig_time.py cap chol_c4 solve_chol_kernel 1 CFG=c4 REPS=2
PROG=tools/krig_time.py cap search_c4 search_kernel 1 CFG=c4 REPS=2
PROG=tools/krig_time.py cap lu_c5 solve_blk_kernel 1 CFG=c5 REPS=2
PROG=tools/krig_time.py cap search_c5 search_kernel 1 CFG=c5 REPS=2
compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02_sanitizer_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r02_sanitizer_racecheck.log
ls -la gpurun_out | head -40
