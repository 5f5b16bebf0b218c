#!/bin/bash
# round 2, GPU batch E: tests; K4 fused bench + profile; c2 with the shared-memory tick Cholesky
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest6.log 2>&1; tail -n 12 gpurun_out/r02_gputest6.log
brief() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.4g e2e %.4g ms/step %.1f frac %.3f bad %d launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['chains_with_error_status'], d['gpu_launches']))
    else: print(l.rstrip()[-300:])
"; }
echo "== c4"; timeout 600 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | brief
echo "== c2"; timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | brief
timeout 600 python scripts/ncu_profile.py c4 > gpurun_out/r02_ncu_c4.log 2>&1; python -c "
import json; d=json.load(open('gpurun_out/r02_ncu_c4.json')); n=d['chains']*d['iterations']
print('k4: ms %.1f dram/step %.0f B, dram GB/s %.0f, issue %.1f fp64 %.1f l2hit %.1f' % (d['duration_ms'], d['dram_bytes']/n, d['dram_bytes']/d['duration_ms']/1e6, d['issue_active_pct'], d['fp64_pipe_pct'], d['l2_hit_pct']))"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>/dev/null | grep -E "k2_|k3_" | awk -F'","' '{print $5, $NF}' | sed 's/"//g' | sort | uniq -c | sort -rn | head -8
