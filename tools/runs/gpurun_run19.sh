set -x
O=gpurun_out
( time python -m pytest tests -m gpu -q -x --durations=3 -k "vof or VOF or enright or zalesak or advect" ) > $O/r2_s19_pytest.log 2>&1; tail -12 $O/r2_s19_pytest.log | cut -c1-300
for w in C2_enright_256_f32 C2_enright_256_f64; do python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu > $O/r2_s19_$w.json 2>> $O/r2_s19.err; IFADV_VOF_KERNEL=lean python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu > $O/r2_s19_${w}_lean.json 2>> $O/r2_s19.err; done
tail -n 5 $O/r2_s19.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_s19_C2*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
    print(f, d['value'], d['ms_per_step'], r['step_frac_of_roofline'], r['ms_per_launch_by_direction'])
PY
