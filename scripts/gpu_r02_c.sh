#!/bin/bash
# round 2, GPU batch C: tests; K1 block-size sweep (time + DRAM bytes); K4 RAM kernel bench + profile
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest4.log 2>&1; tail -n 25 gpurun_out/r02_gputest4.log
out=gpurun_out/r02_abC.txt; : > $out
for t in 256 320 384 448 512; do
  echo "== K1 block=$t" >> $out
  MCMCB_K1_BLOCK=$t python scripts/quick_time.py 1048576 100 2>&1 | grep "N=" | tail -n 2 | cut -c1-150 >> $out
  MCMCB_K1_BLOCK=$t ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k1_step -s 1 -c 1 python scripts/prof_small.py 1048576 100 2>&1 | grep -E "dram__|sm__" >> $out
done
cat $out
for k4 in 0 1; do echo "== c4 K4=$k4"; MCMCB_K4=$k4 timeout 600 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.4g e2e %.4g ms/step %.1f frac %.3f bad %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['chains_with_error_status']))
    else: print(l.rstrip()[-300:])
"; done
timeout 600 python scripts/ncu_profile.py c4 > gpurun_out/r02_ncu_c4.log 2>&1; tail -c 400 gpurun_out/r02_ncu_c4.log; echo
