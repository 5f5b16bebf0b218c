#!/bin/bash
# round 2, GPU batch J: thread-per-chain SCAM kernel (K5): parity tests + C5 bench
timeout 900 python -m pytest tests/test_r02_coverage.py tests/test_k3_scam_parity.py tests/test_pool_diag.py -m gpu -q -k "scam or SCAM or pooled" > gpurun_out/r02_gputest9.log 2>&1; tail -n 30 gpurun_out/r02_gputest9.log
brief() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  value %.4g e2e %.4g ms/step %.1f frac %.3f bad %d tpb %s q %s acc %.3f' % (d['value'], d['e2e']['value'] if 'e2e' in d else 0, d['ms_per_step'], d['roofline']['frac'], d['chains_with_error_status'], d['config'].get('threads_per_block'), d['roofline']['stage2_rate_q'], d['roofline']['accept_rate']))
    elif 'rror' in l: print(l.rstrip()[-300:])
"; }
echo "== c5 K5 (default)"; timeout 600 python bench.py --workload c5 --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | brief
