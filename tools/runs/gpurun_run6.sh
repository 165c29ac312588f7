set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29511 tests/mgpu_check.py > gpurun_out/r2_s6_mgpu8.log 2>&1; grep mgpu_check gpurun_out/r2_s6_mgpu8.log
$TR --master-port 29512 bench.py --gpus 8 --steps 12 --warmup 3 > gpurun_out/r2_s6_bench8.json 2> gpurun_out/r2_s6_bench8.err; tail -c 300 gpurun_out/r2_s6_bench8.err
$TR --master-port 29513 bench.py --gpus 8 --workload C5_strong_2048x1024x512_f32 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2_s6_strong8.json 2> gpurun_out/r2_s6_strong8.err; tail -c 300 gpurun_out/r2_s6_strong8.err
IFADV_SLAB_OVERLAP=0 $TR --master-port 29514 bench.py --gpus 8 --steps 12 --warmup 3 --no-e2e --no-extra > gpurun_out/r2_s6_bench8_noovl.json 2> gpurun_out/r2_s6_bench8_noovl.err
