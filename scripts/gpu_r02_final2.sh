#!/bin/bash
# round 2, second final single-GPU evidence run (after the resident tick / theta-in-shared-memory SCAM kernel):
# full GPU suite, ncu captures keyed by the source hash, bench lines of every workload, launch list of C5
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest_final2.log 2>&1; tail -n 4 gpurun_out/r02_gputest_final2.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1; tail -n 1 gpurun_out/r02_smoke.txt
for w in c3 c2 c4; do timeout 300 python scripts/ncu_profile.py $w > gpurun_out/r02_ncu_$w.log 2>&1 || tail -n 5 gpurun_out/r02_ncu_$w.log; done
timeout 300 python scripts/ncu_profile.py c5 --iters 5 > gpurun_out/r02_ncu_c5.log 2>&1 || tail -n 5 gpurun_out/r02_ncu_c5.log
cp gpurun_out/r02_ncu_c*.json profiles/
timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; tail -c 300 gpurun_out/r02_bench_c3.json; echo
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
timeout 600 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err; tail -c 200 gpurun_out/r02_bench_c5.json; echo
for w in c2 c4 c1; do timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; tail -c 200 gpurun_out/r02_bench_$w.json; echo; done
timeout 300 python bench.py --steps 3 --warmup 3 --dump-stride 10 --no-cpu-baseline > gpurun_out/r02_bench_c3_dump10.json 2> gpurun_out/r02_bench_c3_dump10.err; tail -c 300 gpurun_out/r02_bench_c3_dump10.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_c5.csv python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_launches_c5.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_launches_c2.log 2>&1
