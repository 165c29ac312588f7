O=gpurun_out
timeout 100 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"pois_mult|pois_update_kernel|pois_dir_kernel" -s 9 -c 9 --csv --log-file $O/r2_s69_pois_march_launches.csv python tools/time_poisson.py 512 f32 8 > $O/r2_s69.log 2>&1
grep -v "^==" $O/r2_s69_pois_march_launches.csv | cut -d, -f5,13- | cut -c1-200 | head -50
