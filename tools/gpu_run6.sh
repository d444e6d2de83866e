#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kriging_gpu.py -m gpu -q --timeout 900 -x > gpurun_out/tests_krig.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/tests_krig.log
