set -x
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_general.py::test_named_size_C4_bubble_512_f32 2>&1 | tail -15 > gpurun_out/r2_s5_pytest.log; tail -3 gpurun_out/r2_s5_pytest.log
for v in "" _dr2m1 _dr1m2 _dr1m1; do IFADV_LIB=$PWD/interfaceadvection.jl_b200/libifadv_b200$v.so python bench.py --workload C4_bubble_256_f64 --steps 12 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_s5_f64$v.json 2>> gpurun_out/r2_s5.err; done
for w in C2_enright_256_f32 C2_enright_256_f64 C1_zalesak_128_f64; do python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/r2_s5_$w.json 2>> gpurun_out/r2_s5.err; done
for c in 32 64 128; do IFADV_CHUNK=$c python bench.py --workload C3_dambreak_512x256x256_f32 --steps 12 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_s5_C3_c$c.json 2>> gpurun_out/r2_s5.err; done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ifadv -c 90 --csv --log-file gpurun_out/r2_launches_512.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/r2_s5_ncu_bench.log 2>&1
