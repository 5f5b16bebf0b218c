#!/bin/bash
# One gpurun call: GPU parity tests, the bench lines (ours + reference arm), ncu launch list of the bench command,
# DRAM bytes of one full-size bench launch, ncu --set full of the dominant kernel (reduced steps: ncu replays every
# launch ~40 times), K2/K3 timings.
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log 2>&1
python bench.py > gpurun_out/bench_n1.log 2>&1
python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k1_step -s 3 -c 1 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_launch_dram.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1_step -s 1 -c 1 -f -o gpurun_out/prof_k1 \
    python scripts/prof_small.py 1048576 20 > gpurun_out/prof_k1.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_k1.ncu-rep > gpurun_out/prof_k1_summary.txt 2>&1
python scripts/time_k2.py all > gpurun_out/time_k2.log 2>&1
cat gpurun_out/pytest_gpu.log; tail -1 gpurun_out/bench_n1.log | cut -c1-300; grep -E "dram__|gpu__time" gpurun_out/bench_launch_dram.log; grep "steps=" gpurun_out/time_k2.log | cut -c1-120
