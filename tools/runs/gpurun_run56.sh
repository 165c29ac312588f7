O=gpurun_out
N=${NG:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/time_poisson.py slab > $O/r2_s56_poisson_slab$N.txt 2> $O/r2_s56_slab$N.err; cat $O/r2_s56_poisson_slab$N.txt; tail -3 $O/r2_s56_slab$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 tests/mgpu_check.py > $O/r2_s56_mgpu$N.log 2>&1; grep "mgpu_check" $O/r2_s56_mgpu$N.log | tail -8
