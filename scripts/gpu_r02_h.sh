#!/bin/bash
# round 2, GPU batch H: parity after the K2 matvec change + C2 bench
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gputest7.log 2>&1; tail -n 6 gpurun_out/r02_gputest7.log
brief() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  value %.4g ms/step %.2f frac %.3f bad %d tpb %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['chains_with_error_status'], d['config'].get('threads_per_block')))
    elif 'rror' in l: print(l.rstrip()[-300:])
"; }
for w in 8 16; do echo "== c2 warps=$w"; MCMCB_K2_WARPS=$w timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | brief; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>/dev/null | grep -E "k2_" | awk -F'","' '{print $5, $NF}' | sed 's/"//g' | sort | uniq -c | sort -rn | head -6
