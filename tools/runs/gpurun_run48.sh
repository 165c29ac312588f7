O=gpurun_out
python bench.py --workload C5_strong_2048x1024x512_f32 --steps 5 --warmup 2 --no-e2e --no-cpu > $O/r2_s48_strong1.json 2> $O/r2_s48_strong1.err; tail -3 $O/r2_s48_strong1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_s48_strong1.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['config']['workload'], d['scaling'], round(d['value'],3), round(d['ms_per_step'],3), d['roofline']['step_frac_of_roofline'])
PY
