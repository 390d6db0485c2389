cd /root/repo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r03c_poisson_launches.csv oracle/_ref/dropin_poisson 2000 200000 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[l for l in open('/root/repo/gpurun_out/r03c_poisson_launches.csv') if not l.startswith('==')]
agg=collections.OrderedDict()
for r in csv.DictReader(rows):
    if r.get('Metric Name')!='gpu__time_duration.sum': continue
    v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
    v = v/1000 if u=='ns' else (v*1000 if u=='ms' else v)
    a=agg.setdefault(r['Kernel Name'][:100],[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:16]:
    if 'ub200' in k: print(f"{t/1000:9.2f} ms total {n:5d} launches {t/n:9.1f} us  {k}")
PY
