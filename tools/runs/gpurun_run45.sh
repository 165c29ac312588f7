O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > $O/r2_s45_mgpu.log 2>&1; grep "mgpu_check" $O/r2_s45_mgpu.log | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > $O/r2_s45_bench2.json 2> $O/r2_s45_bench2.err; tail -2 $O/r2_s45_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_s45_bench2.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['config']['workload'], round(d['value'],3), round(d['ms_per_step'],3), d['slab_check'] and d['slab_check'].get('pass'), 'c5', (d.get('c5') or {}).get('value'), (d.get('c5') or {}).get('ms_per_step'))
PY
