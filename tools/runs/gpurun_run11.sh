set -x
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_general.py::test_named_size_C4_bubble_512_f32 2>&1 | tail -25 > gpurun_out/r2_s11_pytest.log; tail -4 gpurun_out/r2_s11_pytest.log
for w in C1_zalesak_128_f64 C2_enright_256_f32 C2_enright_256_f64 C4_bubble_256_f64; do python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_s11_$w.json 2>> gpurun_out/r2_s11.err; done
tail -3 gpurun_out/r2_s11.err
