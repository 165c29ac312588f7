O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --workload C5_strong_2048x1024x512_f32 --steps 10 --warmup 3 --no-e2e > $O/r2_s47_strong8.json 2> $O/r2_s47_strong8.err; tail -2 $O/r2_s47_strong8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_s47_strong8.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['config']['workload'], d['scaling'], round(d['value'],3), round(d['ms_per_step'],3), d['slab_check'] and d['slab_check'].get('pass'))
PY
