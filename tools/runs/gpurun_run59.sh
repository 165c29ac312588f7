O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlpoisson.py -q > $O/r2_s59_pytest_ml.log 2>&1; tail -5 $O/r2_s59_pytest_ml.log
IFADV_ML_BOTTOM=0 timeout 600 python -m pytest tests/test_gpu_mlpoisson.py -q -k "solver or history" > $O/r2_s59_pytest_ml_nobottom.log 2>&1; tail -2 $O/r2_s59_pytest_ml_nobottom.log
for n in 64 128 256 512; do
  timeout 120 python tools/time_mlpoisson.py $n f32 4 >> $O/r2_s59_ml.jsonl 2>> $O/r2_s59_ml.err
  IFADV_ML_BOTTOM=0 timeout 120 python tools/time_mlpoisson.py $n f32 4 >> $O/r2_s59_ml_nobottom.jsonl 2>> $O/r2_s59_ml.err
done
IFADV_ML_BOTTOM=50000 timeout 120 python tools/time_mlpoisson.py 128 f32 4 >> $O/r2_s59_ml_bottom50k.jsonl 2>> $O/r2_s59_ml.err
timeout 120 python tools/time_mlpoisson.py 256 f64 4 >> $O/r2_s59_ml.jsonl 2>> $O/r2_s59_ml.err
python - <<'PY'
import json
for f in ('r2_s59_ml','r2_s59_ml_nobottom','r2_s59_ml_bottom50k'):
    print(f)
    for l in open('gpurun_out/%s.jsonl'%f):
        d=json.loads(l); print(d['grid'][0], d['dtype'], 'ms/cycle %.3f'%d['ms_per_cycle'], 'frac %.3f'%d['frac_of_hbm_roofline'], 'launches %.1f'%d['launches_per_cycle'], d['r2_after_cycles_1_to_4'][-1], d['myproject_to_convergence'])
PY
tail -5 $O/r2_s59_ml.err
