O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_poisson.py tests/test_gpu_mlpoisson.py -q > $O/r2_s63_pytest.log 2>&1; tail -4 $O/r2_s63_pytest.log
timeout 120 python tools/time_poisson.py 512 f32 50 > $O/r2_s63_pois.jsonl 2>> $O/r2_s63.err
timeout 120 python tools/time_poisson.py 256 f32 50 >> $O/r2_s63_pois.jsonl 2>> $O/r2_s63.err
timeout 120 python tools/time_poisson.py 256 float64 50 >> $O/r2_s63_pois.jsonl 2>> $O/r2_s63.err
for n in 64 128 256 512; do timeout 120 python tools/time_mlpoisson.py $n f32 4 >> $O/r2_s63_ml.jsonl 2>> $O/r2_s63.err; done
timeout 120 python tools/time_mlpoisson.py 256 f64 4 >> $O/r2_s63_ml.jsonl 2>> $O/r2_s63.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_s63_pois.jsonl'):
    d=json.loads(l); print('psolver', d['grid'][0], d['dtype'], d['ms_per_iteration'], d['frac_of_hbm_roofline'], d['myproject_to_convergence'])
for l in open('gpurun_out/r2_s63_ml.jsonl'):
    d=json.loads(l); print('ml', d['grid'][0], d['dtype'], d['ms_per_cycle'], d['frac_of_hbm_roofline'], d['launches_per_cycle'], d['myproject_to_convergence'])
PY
tail -3 $O/r2_s63.err
