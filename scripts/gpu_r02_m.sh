#!/bin/bash
brief() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  value %.4g ms/step %.1f frac %.3f bad %d launches %d' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['chains_with_error_status'], d['gpu_launches']))
    elif 'rror' in l: print(l.rstrip()[-300:])
"; }
cp mcmcf90_b200/libmcmcb200.so /tmp/keep.so
echo "== c5 K5 eager minb=4"; timeout 600 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | brief
for mb in 6 8; do cp scratch_libs/k5mb$mb.so mcmcf90_b200/libmcmcb200.so; echo "== c5 K5 eager minb=$mb"; timeout 600 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | brief; done
cp /tmp/keep.so mcmcf90_b200/libmcmcb200.so
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_c5.csv python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_c5.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    agg[r[4][:60]][0]+=1; agg[r[4][:60]][1]+=float(r[-1])/1e6
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:8]: print("%-62s %4d %10.2f ms"%(k,v[0],v[1]))
PY
