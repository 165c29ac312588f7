O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_parity_general.py -q -m gpu -x > $O/r2_s66_pytest.log 2>&1; tail -3 $O/r2_s66_pytest.log
python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu --no-extra > $O/r2_s66_bench.json 2> $O/r2_s66.err; python -c "
import json; d=json.load(open('$O/r2_s66_bench.json')); print('f32', d['value'], d['ms_per_step'], d['roofline']['step_frac_of_roofline'])"
tail -2 $O/r2_s66.err
