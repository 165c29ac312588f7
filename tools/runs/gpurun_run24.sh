set -x
O=gpurun_out
python -m pytest tests -m gpu -q -x -k "host_entry" 2>&1 | tail -3
for cp in 128 64 48 32; do IFADV_HOST_CHUNK=$cp python bench.py --steps 5 --warmup 3 --no-cpu --no-extra --e2e-steps 5 2>>$O/r2_s24.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('chunk $cp share', e['value'], e['ms_per_step'], e['h2d_bytes_per_step'], e['pipeline'][:40])"; done
IFADV_HOST_NOSHARE=1 IFADV_HOST_CHUNK=128 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra --e2e-steps 5 2>>$O/r2_s24.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('chunk 128 noshare', e['value'], e['ms_per_step'], e['h2d_bytes_per_step'])"
tail -n 3 $O/r2_s24.err
