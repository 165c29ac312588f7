O=gpurun_out
( time python -m pytest tests -m gpu -q -x --durations=5 ) > $O/r2_s28_pytest.log 2>&1
python bench.py > $O/r2_s28_bench.json 2> $O/r2_s28_bench.err
for w in C1_zalesak_128_f64 C2_enright_256_f32 C2_enright_256_f64 C3_dambreak_512x256x256_f32 C4_bubble_256_f64; do python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu > $O/r2_s28_$w.json 2>> $O/r2_s28.err; done
tail -6 $O/r2_s28_pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_s28_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('r2_s28_')[1], round(d['value'],3), round(d['ms_per_step'],4), r.get('step_frac_of_roofline'), r.get('frac'), (d.get('e2e') or {}).get('value'), r.get('ms_per_launch_by_direction'))
    except Exception as e: print(f, e)
PY
