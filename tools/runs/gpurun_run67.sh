O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_mlpoisson.py -q > $O/r2_s67_pytest.log 2>&1; tail -3 $O/r2_s67_pytest.log
for a in "512 f32" "256 f32" "256 f64"; do timeout 100 python tools/time_mlpoisson.py $a 4 >> $O/r2_s67_ml.jsonl 2>> $O/r2_s67.err; done
python - <<'PY'
import json
for l in open('gpurun_out/r2_s67_ml.jsonl'):
    d=json.loads(l); print('ml', d['grid'][0], d['dtype'], d['ms_per_cycle'], d['frac_of_hbm_roofline'], d['myproject_to_convergence'])
PY
tail -3 $O/r2_s67.err
