#!/bin/bash
# round 2, last evidence run: pooled parity tests, ncu captures keyed by the final source hash, C4 / C3 bench lines
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_pool_diag.py tests/test_r02_coverage.py -m gpu -q -k "pool or ram" > gpurun_out/r02_gputest_final3.log 2>&1; tail -n 3 gpurun_out/r02_gputest_final3.log
for w in c4 c3 c2; do timeout 150 python scripts/ncu_profile.py $w > gpurun_out/r02_ncu_$w.log 2>&1 || tail -n 5 gpurun_out/r02_ncu_$w.log; done
timeout 150 python scripts/ncu_profile.py c5 --iters 5 > gpurun_out/r02_ncu_c5.log 2>&1 || tail -n 5 gpurun_out/r02_ncu_c5.log
cp gpurun_out/r02_ncu_c*.json profiles/
timeout 120 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err; tail -c 200 gpurun_out/r02_bench_c4.json; echo
timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c3_final3.json 2> gpurun_out/r02_bench_c3_final3.err; tail -c 200 gpurun_out/r02_bench_c3_final3.json; echo
