O=gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests -m gpu -q -x -k "not named_size and not 1000 and not full_size and not zalesak and not revolution and not many_steps and not host_entry and not properties" > $O/r2_s34_memcheck_all.log 2>&1; echo "rc=$?" >> $O/r2_s34_memcheck_all.log
tail -8 $O/r2_s34_memcheck_all.log | cut -c1-300
