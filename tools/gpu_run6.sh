#!/bin/bash
mkdir -p gpurun_out
CFG=c5 REPS=3 python tools/krig_time.py
python tools/quick_krig_lu.py 2>&1 | tail -4
python -m pytest tests/test_kriging_gpu.py -m gpu -q --timeout 900 -x > gpurun_out/tests_krig.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/tests_krig.log
