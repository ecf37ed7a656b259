#!/bin/bash
# compute-sanitizer over the hot path: memcheck and racecheck of the smoke invocation and of
# the tensor-core (tcgen05 / mbarrier) parity tests.  The logs' summaries go to profiles/.
# usage: scripts/sanitizer.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
T="tests/test_gpu_ops.py::test_tensor_core_path_isolated_nodes_and_weight_cache tests/test_gpu_ops.py::test_pipeline_config2_10k tests/test_gpu_ops.py::test_pipeline_split_layout_radius_isolated_nodes_and_frames"
for TOOL in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_sanitizer_${TOOL}_smoke.log 2>&1
  timeout 1800 compute-sanitizer --tool $TOOL --print-limit 20 python -m pytest $T -x -q > $OUT/${TAG}_sanitizer_${TOOL}_tests.log 2>&1
done
for f in $OUT/${TAG}_sanitizer_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|smoke ok|Error" $f | tail -5; done
