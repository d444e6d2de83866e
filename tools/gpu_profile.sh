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
PROG=tools/krig_time.py cap ldl_c5 solve_ldl_kernel 1 CFG=c5 REPS=2
PROG=tools/krig_time.py cap lu_c5 solve_blk_kernel 1 CFG=c5 REPS=2 FLAGS=2
PROG=tools/krig_time.py cap search_c5 search_kernel 1 CFG=c5 REPS=2
# summaries on the box (the reports themselves are ~20 MB each; gpurun carries back <= 64 MiB)
for n in pair_c3 pair_c2 chol_c4 search_c4 ldl_c5 lu_c5 search_c5; do
  python tools/ncu_summary.py gpurun_out/r02_ncu_$n.ncu-rep > gpurun_out/r02_ncu_${n}_summary.txt 2>&1
done
python tools/ncu_lines.py gpurun_out/r02_ncu_pair_c3.ncu-rep pair_kernel 40 > gpurun_out/r02_ncu_pair_c3_lines.txt 2>&1
python tools/ncu_lines.py gpurun_out/r02_ncu_chol_c4.ncu-rep solve_chol 60 > gpurun_out/r02_ncu_chol_c4_lines.txt 2>&1
python tools/ncu_lines.py gpurun_out/r02_ncu_lu_c5.ncu-rep solve_blk 60 > gpurun_out/r02_ncu_lu_c5_lines.txt 2>&1
python tools/ncu_lines.py gpurun_out/r02_ncu_ldl_c5.ncu-rep solve_ldl 70 > gpurun_out/r02_ncu_ldl_c5_lines.txt 2>&1
python tools/ncu_lines.py gpurun_out/r02_ncu_search_c4.ncu-rep search_kernel 50 > gpurun_out/r02_ncu_search_c4_lines.txt 2>&1
python tools/ncu_regions.py gpurun_out/r02_ncu_lu_c5.ncu-rep solve_blk 100 > gpurun_out/r02_ncu_lu_c5_regions.txt 2>&1
mkdir -p profiles
python tools/ncu_counters.py pair_kernel_c3=gpurun_out/r02_ncu_pair_c3.ncu-rep pair_kernel_c2=gpurun_out/r02_ncu_pair_c2.ncu-rep \
   solve_chol_kriging=gpurun_out/r02_ncu_chol_c4.ncu-rep search_kriging=gpurun_out/r02_ncu_search_c4.ncu-rep \
   solve_ldl_kriging_c5=gpurun_out/r02_ncu_ldl_c5.ncu-rep solve_lu_kriging_c5=gpurun_out/r02_ncu_lu_c5.ncu-rep search_kriging_c5=gpurun_out/r02_ncu_search_c5.ncu-rep > gpurun_out/counters.log 2>&1
cp profiles/r02_kernel_counters.json gpurun_out/
rm -f gpurun_out/r02_ncu_ldl_c5.ncu-rep gpurun_out/r02_ncu_lu_c5.ncu-rep gpurun_out/r02_ncu_pair_c3.ncu-rep gpurun_out/r02_ncu_pair_c2.ncu-rep gpurun_out/r02_ncu_search_c4.ncu-rep gpurun_out/r02_ncu_search_c5.ncu-rep gpurun_out/r02_ncu_chol_c4.ncu-rep
du -sh gpurun_out
