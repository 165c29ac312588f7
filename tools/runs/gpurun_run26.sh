O=gpurun_out
python -m pytest tests -m gpu -q -x -k "vof or VOF or enright or zalesak or advect" 2>&1 | tail -2 > $O/r2_s26.txt
for kb in 1 2 4; do for mb in 4 8; do for w in C2_enright_256_f32 C2_enright_256_f64; do
IFADV_VKB=$kb IFADV_VMB=$mb python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu 2>>$O/r2_s26.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$w KB=$kb MB=$mb', round(d['value'],2), round(d['ms_per_step'],4), round(r['step_frac_of_roofline'],3), {k:round(v,4) for k,v in r['ms_per_launch_by_direction'].items()})" >> $O/r2_s26.txt; done; done; done
for ch in 8 32; do IFADV_CHUNK=$ch python bench.py --workload C2_enright_256_f32 --steps 20 --warmup 3 --no-e2e --no-cpu 2>>$O/r2_s26.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('f32 chunk=$ch', round(d['value'],2), round(d['ms_per_step'],4), round(r['step_frac_of_roofline'],3))" >> $O/r2_s26.txt; done
cat $O/r2_s26.txt
