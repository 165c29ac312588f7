O=gpurun_out
timeout 150 python -m pytest tests/test_gpu_mlpoisson.py tests/test_gpu_poisson.py -q > $O/r2_s68_pytest.log 2>&1; tail -3 $O/r2_s68_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
