O=gpurun_out
python -m pytest tests/test_gpu_poisson.py -x -q -m gpu > $O/r2_s53_pytest.log 2>&1; tail -3 $O/r2_s53_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > $O/r2_s53_mgpu.log 2>&1; grep "mgpu_check\|Error\|error" $O/r2_s53_mgpu.log | tail -20
IFADV_SLAB_P2P=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/mgpu_check.py > $O/r2_s53_mgpu_nccl.log 2>&1; grep "projection\|Error\|error" $O/r2_s53_mgpu_nccl.log | tail -8
