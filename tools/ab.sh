#!/bin/bash
# A/B of alternative builds of the library in ONE gpurun call (same box, same clocks).
# usage: tools/ab.sh "<lib1.so> <lib2.so> ..." [columns] [rows]     ("-" = the shipped library)
mkdir -p gpurun_out
for lib in $1; do
  echo "== $lib"
  if [ "$lib" = "-" ]; then
    timeout 600 python tools/perf_probe.py ${2:-1024} ${3:-121} 2>&1 | grep -E "^rep [12]|tuning"
  else
    HB_LIB_PATH=$lib timeout 600 python tools/perf_probe.py ${2:-1024} ${3:-121} 2>&1 | grep -E "^rep [12]|tuning"
  fi
done
