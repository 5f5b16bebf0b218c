#!/bin/bash
# theta-in-shared-memory SCAM kernel, chains per thread x lanes per chain: parity subset + C5 variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_r02_coverage.py tests/test_full_size.py tests/test_k3_scam_parity.py -m gpu -q -k "scam or c5 or pooled" > gpurun_out/r02_gputest15.log 2>&1; tail -n 12 gpurun_out/r02_gputest15.log
one() {
  python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  $1 value %.4g ms/step %.1f frac %.3f bad %d launch_ms %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['chains_with_error_status'], d['roofline']['avg_launch_ms']))
    elif 'rror' in l: print(l.rstrip()[-300:])
"
}
for v in "4 2" "4 1" "8 2" "2 1"; do
  set -- $v
  MCMCB_K5S_LANES=$1 MCMCB_K5S_CPT=$2 timeout 300 python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | one "c5 k5s lanes=$1 cpt=$2"
done
