O=gpurun_out
( time python -m pytest tests/test_gpu_post.py -q -x ) > $O/r2_s31_pytest.log 2>&1; tail -25 $O/r2_s31_pytest.log | cut -c1-300
