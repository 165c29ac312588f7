O=gpurun_out
rm -f $O/r2_s29.txt
for ch in 16 24 32 48 64 128; do
IFADV_CHUNK=$ch python bench.py --workload C3_dambreak_512x256x256_f32 --steps 20 --warmup 3 --no-e2e --no-cpu 2>>$O/r2_s29.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('C3 chunk=$ch', round(d['value'],3), round(d['ms_per_step'],4), round(r['step_frac_of_roofline'],4), {k:round(v,4) for k,v in r['ms_per_launch_by_direction'].items()})" >> $O/r2_s29.txt; done
for ch in 64 96 128 192 256; do
IFADV_CHUNK=$ch python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-extra 2>>$O/r2_s29.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('C4 chunk=$ch', round(d['value'],3), round(d['ms_per_step'],4), round(r['step_frac_of_roofline'],4), {k:round(v,4) for k,v in r['ms_per_launch_by_direction'].items()})" >> $O/r2_s29.txt; done
cat $O/r2_s29.txt
