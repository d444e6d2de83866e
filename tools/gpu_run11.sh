#!/bin/bash
# LDL' panel width variants
for lib in "" build/libs/pw6.so build/libs/pw4.so; do
  echo "== lib=${lib:-default}"
  STIF_LIB=$lib CFG=c5 FLAGS=0 timeout 600 python tools/krig_time.py 2>&1 | tail -1
  STIF_LIB=$lib timeout 600 python tools/quick_krig_lu.py 2>&1 | grep "product_sum k=50 LDL"
done
