O=gpurun_out
( time python -m pytest tests -m gpu -q -x --durations=5 ) > $O/r2_s38_pytest.log 2>&1; tail -5 $O/r2_s38_pytest.log
python bench.py --workload C1_zalesak_128_f64 --steps 200 --warmup 10 --no-e2e --no-cpu > $O/r2_s38_C1.json 2>> $O/r2_s38.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_s38_reference.json 2> $O/r2_s38_reference.err
python bench.py > $O/r2_s38_bench.json 2> $O/r2_s38_bench.err
python - <<'PY'
import json
for f in ['gpurun_out/r2_s38_C1.json','gpurun_out/r2_s38_reference.json','gpurun_out/r2_s38_bench.json']:
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d.get('roofline') or {}
    print(f.split('r2_s38_')[1], round(d['value'],4), round(d['ms_per_step'],4), r.get('step_frac_of_roofline'), r.get('frac'), (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'))
PY
