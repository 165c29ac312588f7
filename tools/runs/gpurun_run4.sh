set -x
python bench.py --steps 12 --warmup 3 > gpurun_out/r2_s4_bench.json 2> gpurun_out/r2_s4_bench.err
tail -c 600 gpurun_out/r2_s4_bench.err
for v in "" _dr2m1 _dr1m2 _dr1m1; do IFADV_LIB=$PWD/interfaceadvection.jl_b200/libifadv_b200$v.so python bench.py --workload C4_bubble_256_f64 --steps 12 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_s4_f64$v.json 2>> gpurun_out/r2_s4_bench.err; done
python bench.py --workload C3_dambreak_512x256x256_f32 --steps 12 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_s4_C3.json 2>> gpurun_out/r2_s4_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 45 -c 40 --csv --log-file gpurun_out/r2_launches_512.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/r2_s4_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:xrow_kernel -s 8 -c 2 -o gpurun_out/r2_xrow_prof python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/r2_s4_ncu_full.log 2>&1
(time python bench.py --impl reference --steps 20 --warmup 3) > gpurun_out/r2_s4_reference.json 2> gpurun_out/r2_s4_reference.err
tail -3 gpurun_out/r2_s4_reference.err
