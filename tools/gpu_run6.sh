#!/bin/bash
mkdir -p gpurun_out
for lib in "" build/libs/prev.so; do echo "=== ${lib:-new}"; STIF_LIB=$lib REPS=3 python tools/krig_time.py; done
python -m pytest tests/test_kriging_gpu.py tests/test_api_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout 900 -x -k "not c2_ and not c3_" > gpurun_out/tests_krig.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/tests_krig.log
