O=gpurun_out
python -m pytest tests/test_gpu_poisson.py -x -q -m gpu > $O/r2_s50_pytest.log 2>&1; tail -15 $O/r2_s50_pytest.log
python tools/time_poisson.py > $O/r2_s50_poisson.txt 2> $O/r2_s50_poisson.err; cat $O/r2_s50_poisson.txt; tail -5 $O/r2_s50_poisson.err
