O=gpurun_out
rm -f $O/r2_s49.txt
for ch in 20 24 28 32 40 48 56; do for w in C2_enright_256_f32 C2_enright_256_f64; do
IFADV_CHUNK=$ch python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu 2>>$O/r2_s49.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$w chunk=$ch', round(d['value'],2), round(d['ms_per_step'],4), round(r['step_frac_of_roofline'],3))" >> $O/r2_s49.txt; done; done
cat $O/r2_s49.txt
