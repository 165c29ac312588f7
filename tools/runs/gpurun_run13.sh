set -x
O=gpurun_out
( time python -m pytest tests -m gpu -q -x --durations=5 ) > $O/r2_s13_pytest.log 2>&1; tail -15 $O/r2_s13_pytest.log
python bench.py --workload C1_zalesak_128_f64 --steps 50 --warmup 5 --no-e2e --no-cpu > $O/r2_s13_C1.json 2> $O/r2_s13.err
tail -n 3 $O/r2_s13.err
