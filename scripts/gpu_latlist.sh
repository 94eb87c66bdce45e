cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
LL_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-330} -c 40 --csv --log-file gpurun_out/lat_launches.csv python scripts/lat_run.py > gpurun_out/lat_launches.log 2>&1
tail -2 gpurun_out/lat_launches.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/lat_launches.csv')) if len(r)>10 and r[0].isdigit()]
tot=0
for r in rows:
    name=r[4].split('(')[0]; v=float(r[-1].replace(',','')); u=r[-2]
    if u=='ns': v/=1000
    elif u=='ms': v*=1000
    tot+=v; print('%-40s %8.2f us  grid %s block %s'%(name[:40], v, r[7] if len(r)>7 else '', r[8] if len(r)>8 else ''))
print('total', tot)
PY
python scripts/time_latency.py
for p in 2 4 8; do echo parts $p; LL_LM_PARTS=$p python scripts/time_latency.py | head -1; done
