#!/bin/bash
# round 2, GPU batch G: C2 warps-per-CTA sweep
brief() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  value %.4g ms/step %.2f frac %.3f bad %d tpb %s blocks %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['chains_with_error_status'], d['config'].get('threads_per_block'), d['config'].get('blocks')))
    elif 'rror' in l: print(l.rstrip()[-300:])
"; }
for w in 8 10 12 14 16; do echo "== c2 warps=$w"; MCMCB_K2_WARPS=$w timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | brief; done
