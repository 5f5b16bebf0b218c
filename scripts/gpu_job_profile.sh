#!/bin/bash
# One gpurun call: GPU parity tests, ncu launch list of the bench command, ncu --set full of the
# dominant kernels (reduced sizes: ncu replays every launch ~40 times), K2/K3 timings.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1_step -s 1 -c 1 -f -o gpurun_out/prof_k1 \
    python scripts/prof_small.py 151552 10 > gpurun_out/prof_k1.log 2>&1
python scripts/time_k2.py all > gpurun_out/time_k2.log 2>&1
N_C2=1024 ncu --set full --clock-control none --import-source on -k regex:k2_step -s 1 -c 1 -f -o gpurun_out/prof_k2_c2 \
    python scripts/time_k2.py c2 > gpurun_out/prof_k2_c2.log 2>&1
cat gpurun_out/pytest_gpu.log gpurun_out/time_k2.log
