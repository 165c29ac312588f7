O=gpurun_out
rm -f $O/r2_s39.txt
run() { env "$@" python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-extra 2>>$O/r2_s39.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$*', round(d['ms_per_step'],4), {k:round(v,4) for k,v in r['ms_per_launch_by_direction'].items()})" >> $O/r2_s39.txt; }
run A=0
run IFADV_CHUNK_Y=64 IFADV_CHUNK_YF=64 IFADV_CHUNK_Z=64 IFADV_CHUNK_ZF=64
run IFADV_CHUNK_Y=48 IFADV_CHUNK_YF=48 IFADV_CHUNK_Z=48 IFADV_CHUNK_ZF=48
run IFADV_CHUNK_Y=32 IFADV_CHUNK_YF=32 IFADV_CHUNK_Z=32 IFADV_CHUNK_ZF=32
run IFADV_CHUNK_Y=88 IFADV_CHUNK_YF=88 IFADV_CHUNK_Z=88 IFADV_CHUNK_ZF=88
run IFADV_CHUNK_Y=104 IFADV_CHUNK_YF=104 IFADV_CHUNK_Z=104 IFADV_CHUNK_ZF=104
run IFADV_CHUNK_X=32 IFADV_CHUNK_XF=32
run IFADV_CHUNK_X=48 IFADV_CHUNK_XF=48
run IFADV_CHUNK_X=24 IFADV_CHUNK_XF=24
run IFADV_CHUNK_X=16 IFADV_CHUNK_XF=16
run IFADV_CHUNK_X=86 IFADV_CHUNK_XF=86
run A=1
cat $O/r2_s39.txt
