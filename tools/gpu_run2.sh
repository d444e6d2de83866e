#!/bin/bash
mkdir -p gpurun_out
for lib in "" build/libs/tj256.so build/libs/old.so; do
  echo "=== lib=${lib:-default(tj128)}"
  STIF_LIB=$lib timeout 300 python tools/vario_time.py 2>&1 | tail -5
done
python -m pytest tests/test_variogram_gpu.py tests/test_fullsize_gpu.py tests/test_api_gpu.py tests/test_kriging_gpu.py -m gpu -q --timeout 900 -x -k "not every_kernel_variant" > gpurun_out/tests.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/tests.log
