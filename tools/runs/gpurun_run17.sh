set -x
O=gpurun_out
( time python -m pytest tests/test_gpu_forcing.py -q -x --durations=3 ) > $O/r2_s17_pytest.log 2>&1; tail -12 $O/r2_s17_pytest.log | cut -c1-300
python tools/time_forcing.py > $O/r2_s17_forcing_512_f32.json 2> $O/r2_s17_forcing.err; tail -n 5 $O/r2_s17_forcing.err; cat $O/r2_s17_forcing_512_f32.json
python tools/time_forcing.py --n 256 --f64 > $O/r2_s17_forcing_256_f64.json 2>> $O/r2_s17_forcing.err; cat $O/r2_s17_forcing_256_f64.json
