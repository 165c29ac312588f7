O=gpurun_out
python bench.py > $O/r2_s44_bench.json 2> $O/r2_s44_bench.err
for w in C1_zalesak_128_f64 C2_enright_256_f32 C2_enright_256_f64 C3_dambreak_512x256x256_f32 C4_bubble_256_f64; do st=20; [ $w = C1_zalesak_128_f64 ] && st=200; python bench.py --workload $w --steps $st --warmup 5 --no-e2e --no-cpu > $O/r2_s44_$w.json 2>> $O/r2_s44.err; done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"xrow_kernel|along2_kernel|axpby_kernel|bcf_kernel|red_init_kernel" -c 60 --csv --log-file $O/r2_s44_launches_512.csv python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-extra > $O/r2_s44_ncu_bench.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_s44_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('r2_s44_')[1], round(d['value'],3), round(d['ms_per_step'],4), r.get('step_frac_of_roofline'), r.get('frac'), (d.get('e2e') or {}).get('value'), (d.get('c5') or {}).get('value'), (d.get('tgv_line') or {}).get('value'))
    except Exception as e: print(f, e)
PY
