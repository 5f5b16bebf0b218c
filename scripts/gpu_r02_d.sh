#!/bin/bash
# round 2, GPU batch D: tests; K4 (unrolled) bench + profile; c3 bench with the 384-thread default
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest5.log 2>&1; tail -n 12 gpurun_out/r02_gputest5.log
for k4 in 1; do echo "== c4 K4=$k4"; MCMCB_K4=$k4 timeout 600 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.4g e2e %.4g ms/step %.1f frac %.3f bad %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['chains_with_error_status']))
    else: print(l.rstrip()[-300:])
"; done
timeout 600 python scripts/ncu_profile.py c4 > gpurun_out/r02_ncu_c4.log 2>&1; python -c "
import json; d=json.load(open('gpurun_out/r02_ncu_c4.json')); n=d['chains']*d['iterations']
print('k4: ms %.1f dram/step %.0f B, dram GB/s %.0f, issue %.1f fp64 %.1f l2hit %.1f' % (d['duration_ms'], d['dram_bytes']/n, d['dram_bytes']/d['duration_ms']/1e6, d['issue_active_pct'], d['fp64_pipe_pct'], d['l2_hit_pct']))"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c3_t384.json 2> gpurun_out/r02_bench_c3_t384.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c3_t384.json')); print('c3 value %.4g e2e %.4g tpb %s' % (d['value'], d['e2e']['value'], d['roofline'].get('profile')))"
