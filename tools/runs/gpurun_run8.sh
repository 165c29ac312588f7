set -x
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_general.py::test_named_size_C4_bubble_512_f32 2>&1 | tail -25 > gpurun_out/r2_s8_pytest.log; tail -4 gpurun_out/r2_s8_pytest.log
python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_s8_arow.json 2> gpurun_out/r2_s8.err
IFADV_KERNEL=along2 python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/r2_s8_along2.json 2>> gpurun_out/r2_s8.err
python bench.py --workload C2_enright_256_f32 --steps 20 --warmup 3 > gpurun_out/r2_s8_C2.json 2>> gpurun_out/r2_s8.err
python bench.py --workload C3_dambreak_512x256x256_f32 --steps 12 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_s8_C3.json 2>> gpurun_out/r2_s8.err
ncu --set full --clock-control none --import-source on -k regex:arow_kernel -s 6 -c 2 -o gpurun_out/r2_arow_prof python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/r2_s8_ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ifadv -c 90 --csv --log-file gpurun_out/r2_launches_512.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/r2_s8_ncu_bench.log 2>&1
