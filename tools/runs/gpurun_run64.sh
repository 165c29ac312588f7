O=gpurun_out
for v in "" _a; do
  lib=$PWD/interfaceadvection.jl_b200/libifadv_b200$v.so
  tag=${v:-_base}
  IFADV_LIB=$lib timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q -k "cmom_sweeps or fused_entry or tiny or golden or families or limiters" 2>&1 | tail -2 > $O/r2_s64_ab$tag.log
  for w in C4_bubble_256_f64 C2_enright_256_f64; do
    IFADV_LIB=$lib timeout 300 python bench.py --workload $w --steps 12 --warmup 3 --no-e2e --no-cpu --no-extra > $O/r2_s64_ab${tag}_$w.json 2>> $O/r2_s64_ab$tag.log
    python - <<PY
import json
try:
    d=json.load(open("$O/r2_s64_ab${tag}_$w.json"))
    print("$tag $w", round(d["value"],3), "Gcell/s", round(d["ms_per_step"],3), "ms", {k:round(v,3) for k,v in d["roofline"]["ms_per_launch_by_direction"].items()})
except Exception as e:
    print("$tag", "FAILED", e)
PY
  done
  tail -2 $O/r2_s64_ab$tag.log
done
