set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
$TR --master-port 29511 tools/mgpu_debug.py > gpurun_out/r2_s7_dbg.log 2>&1; grep dbg gpurun_out/r2_s7_dbg.log | grep -v OK | head -40; grep -c "OK" gpurun_out/r2_s7_dbg.log
for mode in "" "NCCL_MIN_P2P_NCHANNELS=16"; do
env $mode $TR --master-port 29512 bench.py --gpus 4 --workload C5_sloshing_2048x1024x128_f32 --steps 8 --warmup 3 --no-e2e > gpurun_out/r2_s7_c5_4_${mode%%=*}.json 2> gpurun_out/r2_s7_c5_4.err
done
IFADV_SLAB_OVERLAP=0 $TR --master-port 29514 bench.py --gpus 4 --workload C5_sloshing_2048x1024x128_f32 --steps 8 --warmup 3 --no-e2e > gpurun_out/r2_s7_c5_4_noovl.json 2>> gpurun_out/r2_s7_c5_4.err
