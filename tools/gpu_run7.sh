#!/bin/bash
# LDL' kernel: parity tests, then timing against the LU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_kriging_gpu.py -m gpu -q --timeout 900 -x > gpurun_out/tests_krig.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/tests_krig.log
timeout 600 python tools/quick_krig_lu.py 2>&1 | tee gpurun_out/quick_krig_lu.log
CFG=c5 FLAGS=0 timeout 600 python tools/krig_time.py 2>&1 | tee gpurun_out/krig_time_ldl.log
CFG=c5 FLAGS=2 timeout 600 python tools/krig_time.py 2>&1 | tee gpurun_out/krig_time_lu.log
