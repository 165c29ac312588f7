O=gpurun_out
( time python -m pytest tests -q -m gpu ) > $O/r2_s65_pytest_gpu.log 2>&1; tail -6 $O/r2_s65_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_s65_smoke.log 2>&1; tail -3 $O/r2_s65_smoke.log
python bench.py --workload C4_bubble_256_f64 --no-e2e --no-cpu --no-extra > $O/r2_s65_bench_C4_256_f64.json 2> $O/r2_s65_f64.err; python -c "
import json; d=json.load(open('$O/r2_s65_bench_C4_256_f64.json')); print('f64', d['value'], d['ms_per_step'], d['roofline']['step_frac_of_roofline'], d['roofline']['ms_per_launch_by_direction'])"
python bench.py > $O/r2_s65_bench.json 2> $O/r2_s65_bench.err; python -c "
import json; d=json.load(open('$O/r2_s65_bench.json')); print('f32', d['value'], d['ms_per_step'], d['roofline']['step_frac_of_roofline'], d['e2e']['value'], d['projection']['ms_per_iteration'], d['projection']['multigrid']['ms_per_cycle'])"
tail -3 $O/r2_s65_bench.err
