O=gpurun_out
python -m pytest tests/test_gpu_poisson.py -x -q -m gpu > $O/r2_s54_pytest.log 2>&1; tail -3 $O/r2_s54_pytest.log
IFADV_POIS_COOP=0 python -m pytest tests/test_gpu_poisson.py -x -q -m gpu > $O/r2_s54_pytest_nocoop.log 2>&1; tail -3 $O/r2_s54_pytest_nocoop.log
rm -f $O/r2_s54_coop.txt
for n in 32 64 128; do
for v in "IFADV_POIS_COOP=0" "IFADV_POIS_COOP=1" "IFADV_POIS_COOP=1 IFADV_POIS_COOP_CTAS=296" "IFADV_POIS_COOP=1 IFADV_POIS_COOP_CTAS=148"; do
  echo "n=$n $v" >> $O/r2_s54_coop.txt
  env $v python tools/time_poisson.py $n float32 200 2>>$O/r2_s54.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['iterations'], round(d['ms_per_iteration']*1e3,2),'us/iter', d['launches'],'launches', d['myproject_to_convergence'])" >> $O/r2_s54_coop.txt
done; done
env IFADV_POIS_COOP=1 python tools/time_poisson.py 128 float64 200 2>>$O/r2_s54.err | tail -1 >> $O/r2_s54_coop.txt
env IFADV_POIS_COOP=0 python tools/time_poisson.py 128 float64 200 2>>$O/r2_s54.err | tail -1 >> $O/r2_s54_coop.txt
cat $O/r2_s54_coop.txt; tail -3 $O/r2_s54.err
