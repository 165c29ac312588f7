O=gpurun_out
python -m pytest tests/test_gpu_poisson.py -x -q -m gpu > $O/r2_s52_pytest.log 2>&1; tail -3 $O/r2_s52_pytest.log
python tools/time_poisson.py > $O/r2_s52_poisson.txt 2> $O/r2_s52_poisson.err; cat $O/r2_s52_poisson.txt; tail -5 $O/r2_s52_poisson.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pois -c 40 --csv --log-file $O/r2_s52_pois_launches_512.csv python tools/time_poisson.py 512 float32 10 > $O/r2_s52_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2_s52_pois_launches_512.csv') if l.startswith('"'))]
h=rows[0]; d=collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[1:]:
    r=dict(zip(h,r)); d[r['Kernel Name'].split('<')[0].split('(')[0]][r['Metric Name']].append(float(r['Metric Value'].replace(',','')))
for k,m in d.items():
    print(k,{a:(round(sum(v)/len(v),1),m and len(v)) for a,v in m.items()}, [ (a, {r2['Metric Unit'] for r2 in [dict(zip(h,x)) for x in rows[1:]] if r2['Metric Name']==a}) for a in m][:0])
PY
