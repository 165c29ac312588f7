O=gpurun_out
IFADV_POIS_MARCH=24 timeout 600 python -m pytest tests/test_gpu_poisson.py tests/test_gpu_mlpoisson.py -q -x > $O/r2_s61_pytest_march.log 2>&1; tail -3 $O/r2_s61_pytest_march.log
for m in 0 16 32 64 128; do
  IFADV_POIS_MARCH=$m timeout 120 python tools/time_poisson.py 512 f32 50 2>> $O/r2_s61.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('psolver march=$m', d['ms_per_iteration'], d['frac_of_hbm_roofline'])" | tee -a $O/r2_s61_march.txt
done
for m in 0 32 64; do
  IFADV_POIS_MARCH=$m timeout 120 python tools/time_mlpoisson.py 512 f32 4 2>> $O/r2_s61.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('multigrid march=$m', d['ms_per_cycle'], d['frac_of_hbm_roofline'])" | tee -a $O/r2_s61_march.txt
done
IFADV_ML_GRID888=1 timeout 120 python tools/time_mlpoisson.py 512 f32 4 2>> $O/r2_s61.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('multigrid grid888', d['ms_per_cycle'], d['frac_of_hbm_roofline'])" | tee -a $O/r2_s61_march.txt
IFADV_POIS_MARCH=32 timeout 120 python tools/time_poisson.py 256 f64 50 2>> $O/r2_s61.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('psolver 256 f64 march=32', d['ms_per_iteration'], d['frac_of_hbm_roofline'])" | tee -a $O/r2_s61_march.txt
IFADV_POIS_MARCH=0 timeout 120 python tools/time_poisson.py 256 f64 50 2>> $O/r2_s61.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('psolver 256 f64 march=0', d['ms_per_iteration'], d['frac_of_hbm_roofline'])" | tee -a $O/r2_s61_march.txt
tail -3 $O/r2_s61.err
