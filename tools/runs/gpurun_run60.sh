O=gpurun_out
( time python -m pytest tests -q -m gpu ) > $O/r2_s60_pytest_gpu.log 2>&1; tail -6 $O/r2_s60_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_s60_smoke.log 2>&1; tail -8 $O/r2_s60_smoke.log
python bench.py > $O/r2_s60_bench.json 2> $O/r2_s60_bench.err; tail -c 2500 $O/r2_s60_bench.json; tail -3 $O/r2_s60_bench.err
IFADV_ML_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ml_|pois_|perbc" -c 260 --csv --log-file $O/r2_s60_ml_launches_512.csv python tools/time_mlpoisson.py 512 f32 4 > $O/r2_s60_ncu_ml.log 2>&1
IFADV_ML_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"ml_pcg_mult_kernel|ml_pcg_update_kernel|ml_pcg_dir_kernel|ml_increment_kernel|ml_jacobi_kernel|ml_pcg_start_kernel" -s 2 -c 8 -o $O/r2_s60_ml_full -f python tools/time_mlpoisson.py 512 f32 4 > $O/r2_s60_ncu_ml_full.log 2>&1
ncu -i $O/r2_s60_ml_full.ncu-rep --page raw --csv > $O/r2_s60_ml_full_raw.csv 2>/dev/null
ls -la $O/r2_s60_* | head -20
