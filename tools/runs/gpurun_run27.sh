O=gpurun_out
python -m pytest tests -m gpu -q -x -k "vof or VOF or enright or zalesak or advect or nan or divergence" 2>&1 | tail -2 > $O/r2_s27.txt
for ch in 16 32 64 128 256; do for w in C2_enright_256_f32 C2_enright_256_f64; do
IFADV_CHUNK=$ch IFADV_VKB=1 python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu 2>>$O/r2_s27.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$w chunk=$ch KB=1', round(d['value'],2), round(d['ms_per_step'],4), round(r['step_frac_of_roofline'],3), {k:round(v,4) for k,v in r['ms_per_launch_by_direction'].items()})" >> $O/r2_s27.txt; done; done
for ch in 32 64; do for w in C2_enright_256_f32 C2_enright_256_f64; do
IFADV_CHUNK=$ch IFADV_VKB=4 python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu 2>>$O/r2_s27.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$w chunk=$ch KB=4', round(d['value'],2), round(d['ms_per_step'],4), round(r['step_frac_of_roofline'],3))" >> $O/r2_s27.txt; done; done
cat $O/r2_s27.txt
