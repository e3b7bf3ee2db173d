#!/bin/bash
# In-situ cost of the integrator phases (HB_DUP_PROBE build): throughput with
# each phase executed twice.  Usage: tools/dup_probe.sh "0 1 2 4 ..." [cols] [windows]
mkdir -p gpurun_out
for m in $1; do
  echo "== HB_DBG=$m"
  HB_DBG=$m timeout 300 python tools/perf_probe.py ${2:-1024} ${3:-121} 2>&1 | grep -E "^rep [12]"
done
