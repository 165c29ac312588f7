set -x
O=gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct --clock-control none -k regex:"vofcell" -s 12 -c 6 --csv --log-file $O/r2_s20_vofcell.csv python bench.py --workload C2_enright_256_f32 --steps 3 --warmup 3 --no-e2e --no-cpu > $O/r2_s20_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2_s20_vofcell.csv')))
h=None
for r in rows:
    if r and r[0]=='ID': h=r; continue
    if h and len(r)==len(h):
        d=dict(zip(h,r)); print(d['Kernel Name'][:60], d['Metric Name'], d['Metric Value'])
PY
