#!/bin/bash
echo "=== default"; SIZES=100000,1000000 REPS=3 python tools/vario_time.py | tail -2
echo "=== deterministic"; FLAGS=33 SIZES=100000,1000000 REPS=3 python tools/vario_time.py | tail -2
