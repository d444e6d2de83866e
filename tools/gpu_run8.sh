#!/bin/bash
# LDL' kernel: warps-per-target variants
mkdir -p gpurun_out
for w in 0 2 1; do
  echo "== STIF_LDL_WARPS=$w"
  STIF_LDL_WARPS=$w CFG=c5 FLAGS=0 timeout 600 python tools/krig_time.py 2>&1 | tail -1
  STIF_LDL_WARPS=$w timeout 600 python tools/quick_krig_lu.py 2>&1 | grep "product_sum k=50 LDL"
done
