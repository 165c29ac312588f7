#!/bin/bash
# A/B harness for kernel experiments on the GPU box: for every variant library given on the command line (suffixes of
# interfaceadvection.jl_b200/libifadv_b200<suffix>.so; "" = the product build) run a parity subset and a short bench.
#   tools/ab.sh "" _a _b      -> gpurun_out/ab_<suffix>.{json,log}
mkdir -p gpurun_out
for v in "$@"; do
  lib=$PWD/interfaceadvection.jl_b200/libifadv_b200$v.so
  tag=${v:-_base}
  IFADV_LIB=$lib timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q -k "cmom_sweeps or fused_entry or tiny or golden or families or limiters" 2>&1 | tail -3 > gpurun_out/ab$tag.log
  IFADV_LIB=$lib timeout 300 python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu > gpurun_out/ab$tag.json 2>> gpurun_out/ab$tag.log
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab$tag.json"))
    print("$tag", round(d["value"],3), "Gcell/s", round(d["ms_per_step"],3), "ms", {k:round(v,3) for k,v in d["roofline"]["ms_per_launch_by_direction"].items()})
except Exception as e:
    print("$tag", "FAILED", e)
PY
  tail -1 gpurun_out/ab$tag.log
done
