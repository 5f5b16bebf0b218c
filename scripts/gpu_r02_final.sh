#!/bin/bash
# round 2, final single-GPU evidence run: tests, ncu captures keyed by the source hash, bench lines of every workload
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest_final.log 2>&1; tail -n 4 gpurun_out/r02_gputest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1; tail -n 1 gpurun_out/r02_smoke.txt
for w in c3 c2 c4; do timeout 400 python scripts/ncu_profile.py $w > gpurun_out/r02_ncu_$w.log 2>&1 || tail -n 5 gpurun_out/r02_ncu_$w.log; done
timeout 600 python scripts/ncu_profile.py c5 --iters 5 > gpurun_out/r02_ncu_c5.log 2>&1 || tail -n 5 gpurun_out/r02_ncu_c5.log
cp gpurun_out/r02_ncu_c*.json profiles/
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
# --set full of ONE wave of the register kernel (ncu saves / restores device memory per replay pass: a full launch takes too long)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k1_step_kernel -s 1 -c 1 -o gpurun_out/r02_k1_full -f python scripts/ncu_profile.py c3 --child --chains 227328 --iters 20 > gpurun_out/r02_k1_full.log 2>&1
ncu -i gpurun_out/r02_k1_full.ncu-rep --page details > gpurun_out/r02_ncu_k1_full_details.txt 2>&1
ncu -i gpurun_out/r02_k1_full.ncu-rep --page raw --csv > gpurun_out/r02_ncu_k1_full_raw.csv 2>&1
rm -f gpurun_out/r02_k1_full.ncu-rep
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; tail -c 300 gpurun_out/r02_bench_c3.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
timeout 600 python bench.py --steps 5 --warmup 3 --dump-stride 10 --no-cpu-baseline > gpurun_out/r02_bench_c3_dump10.json 2> gpurun_out/r02_bench_c3_dump10.err; tail -c 400 gpurun_out/r02_bench_c3_dump10.json; echo
for w in c2 c4 c5 c1; do timeout 900 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; tail -c 200 gpurun_out/r02_bench_$w.json; echo; done
