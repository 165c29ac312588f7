set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/mgpu_check.py > gpurun_out/r2_s10_mgpu8.log 2>&1; grep -E "mgpu_check" gpurun_out/r2_s10_mgpu8.log | cut -c1-200
timeout 600 $TR --master-port 29512 bench.py --gpus 8 --steps 12 --warmup 3 --no-e2e > gpurun_out/r2_s10_bench8.json 2> gpurun_out/r2_s10_bench8.err; tail -c 300 gpurun_out/r2_s10_bench8.err
timeout 300 $TR --master-port 29513 bench.py --gpus 8 --workload C5_strong_2048x1024x512_f32 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2_s10_strong8.json 2> gpurun_out/r2_s10_strong8.err
