# single-stream evidence: ncu launch list of one HDL-64 scan at one lane (eager launches, so every kernel is listed) and the
# wall-clock latency figures (graph replay) of scripts/time_latency.py
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
LL_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-300} -c 36 --csv --log-file gpurun_out/lat_launches.csv python scripts/lat_run.py > gpurun_out/lat_launches.log 2>&1
python - > gpurun_out/latency_launches.txt <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/lat_launches.csv')) if len(r)>10 and r[0].isdigit()]
tot=0
print("ncu --metrics gpu__time_duration.sum --clock-control none, one lane, HDL-64, LL_GRAPH=0 (36 consecutive launches)")
for r in rows:
    name=r[4].split('(')[0]; v=float(r[-1].replace(',','')); u=r[-2]
    if u=='ns': v/=1000
    elif u=='ms': v*=1000
    tot+=v; print('%-44s %8.2f us  block %-14s grid %s'%(name[:44], v, r[7], r[8]))
print('sum %.1f us' % tot)
PY
python scripts/time_latency.py 2>&1 | tail -4 > gpurun_out/latency.txt
cat gpurun_out/latency_launches.txt | tail -40; cat gpurun_out/latency.txt
