O=gpurun_out
( python -m pytest tests -m gpu -q -x -k "vof or VOF or zalesak or advect or golden or nan or revolution or families or surface" 2>&1 | tail -4 ) > $O/r2_s36.txt
python bench.py --workload C1_zalesak_128_f64 --steps 200 --warmup 10 --no-e2e --no-cpu 2>>$O/r2_s36.err > $O/r2_s36_C1.json; python -c "
import json
d=json.loads(open('$O/r2_s36_C1.json').read().strip().splitlines()[-1]); print('C1', d['value'], d['ms_per_step'], d['gpu_launches'], d['config']['mass_drift_rel'])" >> $O/r2_s36.txt
cat $O/r2_s36.txt; tail -3 $O/r2_s36.err
