set -x
O=gpurun_out
( time python -m pytest tests/test_gpu_forcing.py -q -x --durations=5 ) > $O/r2_s14_pytest.log 2>&1; tail -25 $O/r2_s14_pytest.log
python tools/time_forcing.py > $O/r2_s14_forcing_512_f32.json 2> $O/r2_s14_forcing.err; tail -n 5 $O/r2_s14_forcing.err; cat $O/r2_s14_forcing_512_f32.json
python tools/time_forcing.py --n 256 --f64 > $O/r2_s14_forcing_256_f64.json 2>> $O/r2_s14_forcing.err; cat $O/r2_s14_forcing_256_f64.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"xrow_kernel|along2_kernel|axpby_kernel|bcf_kernel|red_init_kernel" -c 60 --csv --log-file $O/r2_s14_launches_512.csv python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-extra > $O/r2_s14_ncu_bench.log 2>&1
tail -n 3 $O/r2_s14_ncu_bench.log
