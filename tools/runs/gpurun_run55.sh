O=gpurun_out
( time python -m pytest tests -q -m gpu ) > $O/r2_s55_pytest_gpu.log 2>&1; tail -6 $O/r2_s55_pytest_gpu.log
python bench.py > $O/r2_s55_bench.json 2> $O/r2_s55_bench.err; tail -c 1500 $O/r2_s55_bench.json; tail -3 $O/r2_s55_bench.err
python tools/time_poisson.py > $O/r2_s55_poisson.txt 2> $O/r2_s55_poisson.err; cut -c1-60,230-700 $O/r2_s55_poisson.txt; tail -3 $O/r2_s55_poisson.err
python tools/time_poisson.py slab > $O/r2_s55_poisson_slab1.txt 2>> $O/r2_s55_poisson.err; cat $O/r2_s55_poisson_slab1.txt
ncu --set full --clock-control none --import-source on -k regex:"pois_mult_kernel|pois_update_kernel|pois_dir_kernel" -s 6 -c 3 -o $O/r2_s55_pois_full -f python tools/time_poisson.py 512 float32 6 > $O/r2_s55_ncu_full.log 2>&1
ncu -i $O/r2_s55_pois_full.ncu-rep --page raw --csv > $O/r2_s55_pois_full_raw.csv 2>/dev/null
ls -la $O/r2_s55_pois_full* | head
