O=gpurun_out
IFADV_POIS_MARCH=24 timeout 600 python -m pytest tests/test_gpu_poisson.py tests/test_gpu_mlpoisson.py -q -x > $O/r2_s62_pytest_march.log 2>&1; tail -2 $O/r2_s62_pytest_march.log
for m in 0 32 64 512; do
  IFADV_POIS_MARCH=$m timeout 120 python tools/time_poisson.py 512 f32 50 2>> $O/r2_s62.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('psolver 512 f32 march=$m', d['ms_per_iteration'], d['frac_of_hbm_roofline'])" | tee -a $O/r2_s62_march.txt
done
for m in 0 32; do
  IFADV_POIS_MARCH=$m timeout 120 python tools/time_poisson.py 256 float64 50 2>> $O/r2_s62.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('psolver 256', d['dtype'], 'march=$m', d['ms_per_iteration'], d['frac_of_hbm_roofline'])" | tee -a $O/r2_s62_march.txt
  IFADV_POIS_MARCH=$m timeout 120 python tools/time_poisson.py 256 f32 50 2>> $O/r2_s62.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('psolver 256', d['dtype'], 'march=$m', d['ms_per_iteration'], d['frac_of_hbm_roofline'])" | tee -a $O/r2_s62_march.txt
  IFADV_POIS_MARCH=$m timeout 120 python tools/time_mlpoisson.py 512 f32 4 2>> $O/r2_s62.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('multigrid 512 f32 march=$m', d['ms_per_cycle'], d['frac_of_hbm_roofline'])" | tee -a $O/r2_s62_march.txt
done
tail -3 $O/r2_s62.err
