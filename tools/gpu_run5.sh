#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 900 > gpurun_out/tests_multi.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/tests_multi.log
