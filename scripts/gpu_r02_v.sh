#!/bin/bash
# C4: bench line again + launch list of one step (which kernel holds the 12 ms that the step gained?)
mkdir -p gpurun_out
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c4_b.json 2> gpurun_out/r02_bench_c4_b.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_c4_b.json').read().strip().splitlines()[-1]); print('c4', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_launches_c4.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r02_launches_c4.csv")))
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hdr]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
acc = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > mv: acc[r[kn][:60]].append(float(r[mv].replace(",", "")) / 1e6)
for k, v in acc.items():
    if "dfma_peak" not in k: print(k, len(v), "total %.1f ms" % sum(v), "last %.2f" % v[-1])
PY
