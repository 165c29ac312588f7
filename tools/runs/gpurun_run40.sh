O=gpurun_out
rm -f $O/r2_s40.txt
run() { wl=$1; shift; env "$@" python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu --no-extra 2>>$O/r2_s40.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$wl $*', round(d['ms_per_step'],4), {k:round(v,4) for k,v in r['ms_per_launch_by_direction'].items()})" >> $O/r2_s40.txt; }
C4=C4_bubble_512_f32; C3=C3_dambreak_512x256x256_f32
for c in 96 100 108 112 116 120 124; do run $C4 IFADV_CHUNK_Y=$c IFADV_CHUNK_YF=$c IFADV_CHUNK_Z=$c IFADV_CHUNK_ZF=$c; done
for c in 20 28 36 40; do run $C4 IFADV_CHUNK_X=$c IFADV_CHUNK_XF=$c; done
run $C4 IFADV_CHUNK_X=24 IFADV_CHUNK_XF=24 IFADV_CHUNK_Y=104 IFADV_CHUNK_YF=104 IFADV_CHUNK_Z=104 IFADV_CHUNK_ZF=104
for c in 40 44 52 56 60; do run $C3 IFADV_CHUNK_Y=$c IFADV_CHUNK_YF=$c IFADV_CHUNK_Z=$c IFADV_CHUNK_ZF=$c; done
for c in 20 24 28 36; do run $C3 IFADV_CHUNK_Z=$c IFADV_CHUNK_ZF=$c; done
for c in 8 12 20 24; do run $C3 IFADV_CHUNK_X=$c IFADV_CHUNK_XF=$c; done
cat $O/r2_s40.txt
