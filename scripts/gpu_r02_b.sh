#!/bin/bash
# round 2, GPU batch B: tests, restart debug, super-tile A/B, ncu profile + bench of every workload
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest3.log 2>&1; tail -n 15 gpurun_out/r02_gputest3.log
python scripts/debug_restart.py > gpurun_out/r02_debug_restart.txt 2>&1
out=gpurun_out/r02_abB.txt; : > $out
for st in 0 1; do
  echo "== supertile=$st" >> $out
  MCMCB_K1_SUPERTILE=$st python scripts/quick_time.py 1048576 100 2>&1 | grep "N=" | tail -n 2 | cut -c1-150 >> $out
done
cat $out
for w in c3 c2 c4; do timeout 600 python scripts/ncu_profile.py $w > gpurun_out/r02_ncu_$w.log 2>&1; tail -c 300 gpurun_out/r02_ncu_$w.log; echo; done
timeout 900 python scripts/ncu_profile.py c5 --iters 5 > gpurun_out/r02_ncu_c5.log 2>&1; tail -c 300 gpurun_out/r02_ncu_c5.log; echo
cp gpurun_out/r02_ncu_c*.json profiles/ 2>/dev/null
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; tail -c 400 gpurun_out/r02_bench_c3.json; tail -n 3 gpurun_out/r02_bench_c3.err
for w in c2 c4 c5 c1; do timeout 900 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; tail -c 300 gpurun_out/r02_bench_$w.json; tail -n 3 gpurun_out/r02_bench_$w.err; done
