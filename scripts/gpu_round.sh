#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (headline + config 2), ncu launch list and a full capture of the named kernels.
# usage: scripts/gpu_round.sh <tag> [kernel-regex ...]
set -u
TAG=${1:-r01}; shift || true
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $OUT/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python bench.py --steps 20 --warmup 5 --workload config2_10k_k16_4x64 --no-cpu-baseline > $OUT/${TAG}_bench_config2.json 2>> $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > $OUT/${TAG}_ncu_launches.log 2>&1
for K in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 8 -c 2 -f -o $OUT/${TAG}_prof_$K \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > $OUT/${TAG}_ncu_$K.log 2>&1
done
cat $OUT/${TAG}_tests.log $OUT/${TAG}_smoke.log $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
