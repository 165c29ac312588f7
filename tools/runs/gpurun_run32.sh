O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_s32_smoke.log 2>&1; tail -6 $O/r2_s32_smoke.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > $O/r2_s32_mgpu.log 2>&1; tail -5 $O/r2_s32_mgpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2_s32_bench2.json 2> $O/r2_s32_bench2.err; tail -3 $O/r2_s32_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_s32_bench2.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['config']['workload'], d['value'], d['ms_per_step'], d['slab_check'] and d['slab_check'].get('pass'), (d.get('c5') or {}).get('value'), (d.get('e2e') or {}).get('value'))
PY
