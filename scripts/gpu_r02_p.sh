#!/bin/bash
# resident tick: parity subset, C5 / C2 benches with the tick resident and streamed, launch list of one C5 step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_r02_coverage.py tests/test_full_size.py tests/test_k2_parity.py tests/test_k3_scam_parity.py tests/test_pool_diag.py -m gpu -q -x > gpurun_out/r02_gputest12.log 2>&1; tail -n 6 gpurun_out/r02_gputest12.log
one() {
  python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  $1 value %.4g ms/step %.1f frac %.3f bad %d launches %d' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['chains_with_error_status'], d['gpu_launches']))
    elif 'rror' in l: print(l.rstrip()[-300:])
"
}
for r in 1 0; do
  MCMCB_TICK_RESIDENT=$r timeout 300 python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | one "c5 resident=$r"
  MCMCB_TICK_RESIDENT=$r timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | one "c2 resident=$r"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_c5_p.csv python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_launches_c5_p.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_c2_p.csv python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_launches_c2_p.log 2>&1
python - <<'PY'
import csv, collections
for w in ("c5", "c2"):
    rows = list(csv.reader(open("gpurun_out/r02_launches_%s_p.csv" % w)))
    hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hdr]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    acc = collections.defaultdict(list)
    for r in rows[hdr + 1:]:
        if len(r) > mv: acc[r[kn][:50]].append(float(r[mv].replace(",", "")) / 1e6)
    for k, v in acc.items():
        if "dfma_peak" not in k: print(w, k, len(v), "total %.1f ms" % sum(v), "last %.2f" % v[-1])
PY
