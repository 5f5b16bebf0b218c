#!/bin/bash
# lane-split theta-in-shared-memory SCAM kernel, tiled pooled covariance: parity subset, C5 variants, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_r02_coverage.py tests/test_full_size.py tests/test_k3_scam_parity.py tests/test_pool_diag.py tests/test_multiproc.py -m gpu -q > gpurun_out/r02_gputest14.log 2>&1; tail -n 12 gpurun_out/r02_gputest14.log
one() {
  python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  $1 value %.4g ms/step %.1f frac %.3f bad %d launches %d' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['chains_with_error_status'], d['gpu_launches']))
    elif 'rror' in l: print(l.rstrip()[-300:])
"
}
for L in 4 2 1; do
  MCMCB_K5S_LANES=$L timeout 300 python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | one "c5 k5s lanes=$L"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_c5_r.csv python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_launches_c5_r.log 2>&1
python - <<'PY'
import csv, collections
for w in ("c5",):
    rows = list(csv.reader(open("gpurun_out/r02_launches_%s_r.csv" % w)))
    hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hdr]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    acc = collections.defaultdict(list)
    for r in rows[hdr + 1:]:
        if len(r) > mv: acc[r[kn][:50]].append(float(r[mv].replace(",", "")) / 1e6)
    for k, v in acc.items():
        if "dfma_peak" not in k: print(w, k, len(v), "total %.1f ms" % sum(v), "last %.2f" % v[-1])
PY
