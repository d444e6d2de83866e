#!/bin/bash
for lib in "" build/libs/minb3.so build/libs/minb5.so; do echo "=== ${lib:-new}"; STIF_LIB=$lib SIZES=100000,300000,1000000 REPS=3 python tools/vario_time.py | tail -3; done
