#!/bin/bash
for lib in "" build/libs/s1_q512.so build/libs/s2_q256.so; do
  echo "=== lib=${lib:-default}"
  SIZES=100000,300000,1000000 REPS=3 STIF_LIB=$lib timeout 300 python tools/vario_time.py 2>&1 | tail -3
done
python -m pytest tests/test_variogram_gpu.py tests/test_fullsize_gpu.py tests/test_api_gpu.py -m gpu -q --timeout 900 > gpurun_out/tests.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/tests.log
