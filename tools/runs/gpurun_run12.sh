set -x
O=gpurun_out
( time python -m pytest tests -m gpu -q -x --durations=8 ) > $O/r2_s12_pytest.log 2>&1; tail -15 $O/r2_s12_pytest.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_s12_reference.json 2> $O/r2_s12_reference.err
python bench.py > $O/r2_s12_bench.json 2> $O/r2_s12_bench.err
for w in C1_zalesak_128_f64 C2_enright_256_f32 C2_enright_256_f64; do python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu > $O/r2_s12_$w.json 2>> $O/r2_s12.err; done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/r2_s12_launches_512.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extra > $O/r2_s12_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"xrow_kernel|along2_kernel" -s 18 -c 3 -o $O/r2_s12_sweeps_prof python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extra > $O/r2_s12_ncu_full.log 2>&1
ncu -i $O/r2_s12_sweeps_prof.ncu-rep --page raw --csv > $O/r2_s12_sweeps_full_raw.csv 2>/dev/null
tail -3 $O/r2_s12.err $O/r2_s12_bench.err
