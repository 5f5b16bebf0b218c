#!/bin/bash
# round 2, GPU batch I: CTA-size choice of the register kernel at small populations (strong-scaling shards)
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gputest8.log 2>&1; tail -n 4 gpurun_out/r02_gputest8.log
brief() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  value %.4g ms/step %.2f frac %.3f bad %d tpb %s blocks %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['chains_with_error_status'], d['config'].get('threads_per_block'), d['config'].get('blocks')))
    elif 'rror' in l: print(l.rstrip()[-300:])
"; }
for n in 1048576 524288 262144 131072; do
  echo "== c3 chains=$n auto"; timeout 300 python bench.py --chains $n --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | brief
  echo "== c3 chains=$n block=384"; MCMCB_K1_BLOCK=384 timeout 300 python bench.py --chains $n --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | brief
done
