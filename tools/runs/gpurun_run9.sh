set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/mgpu_check.py > gpurun_out/r2_s9_mgpu2.log 2>&1; grep -E "mgpu_check|rror" gpurun_out/r2_s9_mgpu2.log | head
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --workload C5_sloshing_2048x1024x128_f32 --steps 8 --warmup 3 --no-e2e > gpurun_out/r2_s9_c5_p2p.json 2> gpurun_out/r2_s9.err
IFADV_SLAB_P2P=0 timeout 300 $TR --master-port 29513 bench.py --gpus 2 --workload C5_sloshing_2048x1024x128_f32 --steps 8 --warmup 3 --no-e2e > gpurun_out/r2_s9_c5_nccl.json 2>> gpurun_out/r2_s9.err
IFADV_SLAB_OVERLAP=1 timeout 300 $TR --master-port 29514 bench.py --gpus 2 --workload C5_sloshing_2048x1024x128_f32 --steps 8 --warmup 3 --no-e2e > gpurun_out/r2_s9_c5_p2p_ovl.json 2>> gpurun_out/r2_s9.err
tail -5 gpurun_out/r2_s9.err
