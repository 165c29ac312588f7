O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlpoisson.py -q > $O/r2_s58_pytest_ml.log 2>&1; tail -25 $O/r2_s58_pytest_ml.log
for n in 64 128 256 512; do
  timeout 120 python tools/time_mlpoisson.py $n f32 4 >> $O/r2_s58_ml.jsonl 2>> $O/r2_s58_ml.err
done
timeout 120 python tools/time_mlpoisson.py 256 f64 4 >> $O/r2_s58_ml.jsonl 2>> $O/r2_s58_ml.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_s58_ml.jsonl'):
    d=json.loads(l); print(d['grid'][0], d['dtype'], 'ms/cycle %.3f'%d['ms_per_cycle'], 'frac %.3f'%d['frac_of_hbm_roofline'], d['r2_after_cycles_1_to_4'], d['myproject_to_convergence'])
PY
tail -5 $O/r2_s58_ml.err
