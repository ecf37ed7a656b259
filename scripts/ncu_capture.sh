#!/bin/bash
# Full ncu capture of the named kernels (regex each) from a short bench run.
# usage: scripts/ncu_capture.sh <tag> <kernel-regex> [...]
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for K in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 8 -c 2 -f -o $OUT/${TAG}_prof_$K \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_$K.log 2>&1
done
ls -la $OUT
