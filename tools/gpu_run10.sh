#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_kriging_gpu.py tests/test_api_gpu.py -m gpu -q --timeout 900 -x > gpurun_out/tests_krig.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/tests_krig.log
echo "== new"; CFG=c4 REPS=4 timeout 600 python tools/krig_time.py 2>&1 | tail -1
echo "== prev"; STIF_LIB=build/libs/prev.so CFG=c4 REPS=4 timeout 600 python tools/krig_time.py 2>&1 | tail -1
