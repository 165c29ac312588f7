O=gpurun_out
python -m pytest tests -m gpu -q -x -k "host_entry" 2>&1 | tail -3 > $O/r2_s25_e2e.txt
for cp in 128 96 64; do IFADV_HOST_CHUNK=$cp python bench.py --steps 5 --warmup 3 --no-cpu --no-extra --e2e-steps 6 2>>$O/r2_s25.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('chunk $cp share', e['value'], e['ms_per_step'], e['h2d_bytes_per_step'], e['pipeline'][:12])" >> $O/r2_s25_e2e.txt; done
cat $O/r2_s25_e2e.txt
