#!/bin/bash
# ncu --set full capture of the one-warp LDL' kernel (product_sum, k = 50)
mkdir -p gpurun_out
env NT=200000 ncu --set full --import-source on --clock-control none -k "regex:solve_ldl_kernel" -s 2 -c 1 -f \
    -o gpurun_out/ncu_ldl_k50 python tools/quick_krig_lu.py > gpurun_out/ncu_ldl_k50.log 2>&1; echo "rc=$? $(tail -1 gpurun_out/ncu_ldl_k50.log)"
python tools/ncu_summary.py gpurun_out/ncu_ldl_k50.ncu-rep > gpurun_out/ncu_ldl_k50_summary.txt 2>&1
python tools/ncu_lines.py gpurun_out/ncu_ldl_k50.ncu-rep solve_ldl 70 > gpurun_out/ncu_ldl_k50_lines.txt 2>&1
rm -f gpurun_out/ncu_ldl_k50.ncu-rep
cat gpurun_out/ncu_ldl_k50_summary.txt | head -40
