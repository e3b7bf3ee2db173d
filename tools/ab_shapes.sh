#!/bin/bash
# A/B of the experimental shapes of one library build (HB_SHAPES=1) in ONE gpurun call.
# usage: tools/ab_shapes.sh <lib.so> "<shape> <shape> ..." [columns] [rows]
for s in $2; do
  echo "== shape $s"
  HB_LIB_PATH=$1 HB_SHAPE=$s timeout 600 python tools/perf_probe.py ${3:-2368} ${4:-61} 2>&1 | grep -E "^rep [12]|tuning|wtd_est|rror|phase|cycles"
done
