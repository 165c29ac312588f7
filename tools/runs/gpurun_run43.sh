O=gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"along2_kernel" -s 14 -c 4 -o /tmp/r2_s43_c3 python bench.py --workload C3_dambreak_512x256x256_f32 --steps 3 --warmup 3 --no-e2e --no-cpu --no-extra > $O/r2_s43_ncu.log 2>&1
ncu -i /tmp/r2_s43_c3.ncu-rep --page raw --csv > $O/r2_s43_c3_raw.csv 2>/dev/null
ncu -i /tmp/r2_s43_c3.ncu-rep --page source --csv > /tmp/r2_s43_c3_src.csv 2>/dev/null
for k in 0 1 2 3; do python tools/ncu_source_hist.py /tmp/r2_s43_c3_src.csv $k > $O/r2_s43_hist_$k.txt 2>&1; done
ls -la $O/r2_s43*
