#!/bin/bash
echo "== new"; CFG=c4,c5 REPS=4 timeout 600 python tools/krig_time.py 2>&1 | tail -2
echo "== prev"; STIF_LIB=build/libs/prev.so CFG=c4,c5 REPS=4 timeout 600 python tools/krig_time.py 2>&1 | tail -2
timeout 1500 python -m pytest tests/test_kriging_gpu.py tests/test_api_gpu.py -m gpu -q --timeout 900 -x 2>&1 | tail -2
