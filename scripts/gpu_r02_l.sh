#!/bin/bash
timeout 900 python -m pytest tests/test_r02_coverage.py tests/test_k3_scam_parity.py tests/test_pool_diag.py -m gpu -q -k "scam or SCAM or pooled" > gpurun_out/r02_gputest10.log 2>&1; tail -n 30 gpurun_out/r02_gputest10.log
brief() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  value %.4g e2e %.4g ms/step %.1f frac %.3f bad %d tpb %s launches %d' % (d['value'], d['e2e']['value'] if 'e2e' in d else 0, d['ms_per_step'], d['roofline']['frac'], d['chains_with_error_status'], d['config'].get('threads_per_block'), d['gpu_launches']))
    elif 'rror' in l: print(l.rstrip()[-300:])
"; }
echo "== c5 K5 lazy view"; timeout 600 python bench.py --workload c5 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | brief
timeout 900 python scripts/ncu_profile.py c5 --iters 5 > gpurun_out/r02_ncu_c5.log 2>&1 || tail -n 5 gpurun_out/r02_ncu_c5.log
python -c "
import json; d=json.load(open('gpurun_out/r02_ncu_c5.json')); n=d['chains']*d['iterations']
print(d['kernel']); print('k5: ms %.1f dram/sweep %.0f B, dram GB/s %.0f, issue %.1f fp64 %.1f l2hit %.1f warps %.1f inst/comp %.0f' % (d['duration_ms'], d['dram_bytes']/n, d['dram_bytes']/d['duration_ms']/1e6, d['issue_active_pct'], d['fp64_pipe_pct'], d['l2_hit_pct'], d['warps_active_pct'], d['warp_inst_executed']*32/n/200))"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>/dev/null | grep -E "k[0-9]_|pool|diag" | awk -F'","' '{print $5, $NF}' | sed 's/"//g' | cut -c1-120
