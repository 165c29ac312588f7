set -x
O=gpurun_out
( time python -m pytest tests -m gpu -q -x -k "vof or VOF or enright or zalesak or advect" ) > $O/r2_s22_pytest.log 2>&1; tail -4 $O/r2_s22_pytest.log | cut -c1-300
for w in C2_enright_256_f32 C2_enright_256_f64; do python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu > $O/r2_s22_$w.json 2>> $O/r2_s22.err; done
tail -n 5 $O/r2_s22.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"vofcell" -s 12 -c 6 --csv --log-file $O/r2_s22_vofcell.csv python bench.py --workload C2_enright_256_f32 --steps 3 --warmup 3 --no-e2e --no-cpu > $O/r2_s22_ncu.log 2>&1
python - <<'PY'
import json,glob,csv
for f in sorted(glob.glob('gpurun_out/r2_s22_C2*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
    print(f, d['value'], d['ms_per_step'], r['step_frac_of_roofline'], r['ms_per_launch_by_direction'])
rows=list(csv.reader(open('gpurun_out/r2_s22_vofcell.csv')))
h=None
for r in rows:
    if r and r[0]=='ID': h=r; continue
    if h and len(r)==len(h):
        d=dict(zip(h,r)); print(d['Kernel Name'][:40], d['Metric Name'], d['Metric Value'])
PY
