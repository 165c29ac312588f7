O=gpurun_out
rm -f $O/r2_s42.txt
run() { wl=$1; shift; env "$@" python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu --no-extra 2>>$O/r2_s42.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$wl $*', round(d['value'],3), round(d['ms_per_step'],4), round(r['step_frac_of_roofline'],4), {k:round(v,4) for k,v in r['ms_per_launch_by_direction'].items()})" >> $O/r2_s42.txt; }
run C4_bubble_512_f32 A=0
run C3_dambreak_512x256x256_f32 A=0
run C5_sloshing_2048x1024x128_f32 A=0
run C4_bubble_256_f64 A=0
run C4_bubble_512_f32_tgv A=0
cat $O/r2_s42.txt
( python -m pytest tests -m gpu -q -x 2>&1 | tail -3 ) >> $O/r2_s42.txt
tail -3 $O/r2_s42.txt
