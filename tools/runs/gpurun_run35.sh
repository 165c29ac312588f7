O=gpurun_out
python -m pytest tests -m gpu -q -x -k "vof or VOF or enright or zalesak or advect or nan or divergence" 2>&1 | tail -2 > $O/r2_s35.txt
for w in C2_enright_256_f32 C2_enright_256_f64; do
python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu 2>>$O/r2_s35.err > $O/r2_s35_$w.json; python -c "
import sys,json
d=json.loads(open('$O/r2_s35_$w.json').read().strip().splitlines()[-1]); r=d['roofline']; print('$w', round(d['value'],2), round(d['ms_per_step'],4), round(r['step_frac_of_roofline'],3), round(r['frac'],3), {k:round(v,4) for k,v in r['ms_per_launch_by_direction'].items()})" >> $O/r2_s35.txt; done
IFADV_CHUNK=8 python -m pytest tests/test_gpu_parity_general.py -m gpu -q -x -k "pure_vof_cell" 2>&1 | tail -1 >> $O/r2_s35.txt
cat $O/r2_s35.txt
