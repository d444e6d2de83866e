#!/bin/bash
mkdir -p gpurun_out
SIZES=100000 REPS=2 ncu --set full --import-source on --clock-control none -k regex:pair_kernel -s 1 -c 1 \
     -o gpurun_out/pair_fx1_c2 -f python tools/vario_time.py > gpurun_out/ncu_fx1.log 2>&1
tail -2 gpurun_out/ncu_fx1.log
