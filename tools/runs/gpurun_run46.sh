O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e > $O/r2_s46_bench8.json 2> $O/r2_s46_bench8.err; tail -2 $O/r2_s46_bench8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_s46_bench8.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['config']['workload'], round(d['value'],3), round(d['ms_per_step'],3), d['slab_check'] and d['slab_check'].get('pass'), 'c5', (d.get('c5') or {}).get('value'), (d.get('c5') or {}).get('ms_per_step'), d['config'].get('parallelism','')[-60:])
PY
