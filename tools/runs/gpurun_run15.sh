set -x
O=gpurun_out
( time python -m pytest tests/test_gpu_forcing.py -q -x --durations=5 ) > $O/r2_s15_pytest.log 2>&1; tail -25 $O/r2_s15_pytest.log | cut -c1-300
