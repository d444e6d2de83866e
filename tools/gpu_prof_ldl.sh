#!/bin/bash
# ncu --set full capture of the LDL' kernel (C5 sample) with per-line summary
mkdir -p gpurun_out
NT=${NT:-40000}
env CFG=c5 NT=$NT REPS=2 ncu --set full --import-source on --clock-control none -k "regex:solve_ldl_kernel" -s 1 -c 1 -f \
    -o gpurun_out/ncu_ldl_c5 python tools/krig_time.py > gpurun_out/ncu_ldl_c5.log 2>&1; echo "rc=$? $(tail -1 gpurun_out/ncu_ldl_c5.log)"
python tools/ncu_summary.py gpurun_out/ncu_ldl_c5.ncu-rep > gpurun_out/ncu_ldl_c5_summary.txt 2>&1
python tools/ncu_lines.py gpurun_out/ncu_ldl_c5.ncu-rep solve_ldl 70 > gpurun_out/ncu_ldl_c5_lines.txt 2>&1
rm -f gpurun_out/ncu_ldl_c5.ncu-rep
cat gpurun_out/ncu_ldl_c5_summary.txt | head -60
