#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench lines of every BASELINE config, ncu launch list and full
# captures of the named kernels.
# usage: scripts/gpu_round.sh <tag> [kernel-regex ...]          (env: SKIP_TESTS=1, CONFIGS="name ...")
set -u
TAG=${1:-r02}; shift || true
OUT=gpurun_out
mkdir -p $OUT
if [ -z "${SKIP_TESTS:-}" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $OUT/${TAG}_tests.log
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_headline.json 2> $OUT/${TAG}_bench.err
for W in ${CONFIGS-config1_300_r3_1x64 config2_10k_k16_4x64 config3_64x300_k20_8x128 config4_64x2000_k20_ppf_4x64 config5_125k_k16_4x64}; do
  timeout 600 python bench.py --steps 20 --warmup 5 --workload $W --no-cpu-baseline > $OUT/${TAG}_bench_$W.json 2>> $OUT/${TAG}_bench.err
done
if [ -n "${REFERENCE:-}" ]; then
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
fi
if [ -n "${LAUNCHES:-}" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > $OUT/${TAG}_ncu_launches.log 2>&1
fi
for K in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 8 -c 2 -f -o $OUT/${TAG}_prof_$K \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > $OUT/${TAG}_ncu_$K.log 2>&1
done
[ -z "${SKIP_TESTS:-}" ] && cat $OUT/${TAG}_tests.log
cat $OUT/${TAG}_smoke.log; for f in $OUT/${TAG}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["config"]["workload"], "ms/step", round(d["ms_per_step"], 4), "edges/s %.3g" % d["value"],
          "e2e ms", round(d["e2e"]["ms_per_step"], 3), "path frac", round(d.get("path_roofline", {}).get("frac", 0), 4),
          {k: round(v, 4) for k, v in list(d.get("kernel_ms_per_step", {}).items())[:7]})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done; tail -5 $OUT/${TAG}_bench.err
