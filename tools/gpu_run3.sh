#!/bin/bash
for lib in "" build/libs/prev.so; do echo "=== ${lib:-new}"; STIF_LIB=$lib SIZES=100000,300000,1000000 REPS=3 python tools/vario_time.py | tail -3; done
python -m pytest tests/test_variogram_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout 900 -x -k "not c4 and not c5 and not pipelined" > gpurun_out/tests_v.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/tests_v.log
