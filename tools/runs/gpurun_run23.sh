set -x
O=gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"vofcell_kernel" -s 6 -c 3 -o $O/r2_s23_vofcell python bench.py --workload C2_enright_256_f32 --steps 3 --warmup 3 --no-e2e --no-cpu > $O/r2_s23_ncu.log 2>&1
ncu -i $O/r2_s23_vofcell.ncu-rep --page raw --csv > $O/r2_s23_vofcell_raw.csv 2>/dev/null
ncu -i $O/r2_s23_vofcell.ncu-rep --page details 2>/dev/null | grep -v "^ *$" | head -150 > $O/r2_s23_details.txt
