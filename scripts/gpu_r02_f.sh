#!/bin/bash
# round 2, GPU batch F (1 GPU): K4 CTA size / unroll variants on C4
mkdir -p gpurun_out
brief() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.4g e2e %.4g ms/step %.1f frac %.3f bad %d tpb %s blocks %s' % (d['value'], d['e2e']['value'] if 'e2e' in d else 0, d['ms_per_step'], d['roofline']['frac'], d['chains_with_error_status'], d['config'].get('threads_per_block'), d['config'].get('blocks')))
    else: print(l.rstrip()[-300:])
"; }
cp mcmcf90_b200/libmcmcb200.so /tmp/keep.so
for b in auto 128 64 32; do
  echo "== c4 U=8 block=$b"; if [ $b = auto ]; then unset MCMCB_K4_BLOCK; else export MCMCB_K4_BLOCK=$b; fi
  timeout 600 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | brief
done
unset MCMCB_K4_BLOCK
cp scratch_libs/k4u16.so mcmcf90_b200/libmcmcb200.so
echo "== c4 U=16 block=auto"; timeout 600 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | brief
cp /tmp/keep.so mcmcf90_b200/libmcmcb200.so
echo "== c2 (tick Cholesky back in L2)"; timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | brief
