O=gpurun_out
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_forcing.py tests/test_gpu_post.py -q -x -k "not 140" > $O/r2_s33_memcheck_forcing.log 2>&1; echo "rc=$?" >> $O/r2_s33_memcheck_forcing.log
tail -8 $O/r2_s33_memcheck_forcing.log | cut -c1-300
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity_general.py -q -x -k "pure_vof_cell or exit_bc or three_distinct" > $O/r2_s33_memcheck_vof.log 2>&1; echo "rc=$?" >> $O/r2_s33_memcheck_vof.log
tail -8 $O/r2_s33_memcheck_vof.log | cut -c1-300
