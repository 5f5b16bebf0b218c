#!/bin/bash
# --set full of ONE wave of the SCAM kernel with theta in shared memory, 4 lanes x 2 chains per thread; clocks of a bench run
mkdir -p gpurun_out
MCMCB_K5S_LANES=4 MCMCB_K5S_CPT=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k5s_scam_step_kernel -s 1 -c 1 -o gpurun_out/r02_k5s_full_L4C2 -f python scripts/ncu_profile.py c5 --child --chains 18944 --iters 3 > gpurun_out/r02_k5s_full_L4C2.log 2>&1
ncu -i gpurun_out/r02_k5s_full_L4C2.ncu-rep --page details > gpurun_out/r02_ncu_k5s_full_L4C2_details.txt 2>&1
ncu -i gpurun_out/r02_k5s_full_L4C2.ncu-rep --page source --csv > gpurun_out/r02_ncu_k5s_full_L4C2_source.csv 2>&1
rm -f gpurun_out/r02_k5s_full_L4C2.ncu-rep
grep -E "Duration|Executed Ipc Active|Issue Slots Busy|No Eligible|Eligible Warps|Achieved Occupancy|highest-utilized|Registers Per|L1/TEX Hit|Mem Busy|Mem Pipes|way bank" gpurun_out/r02_ncu_k5s_full_L4C2_details.txt | head -30
timeout 300 python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_c5_L4C2.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c5_L4C2.json')); print(d['value'], d['ms_per_step'], d['clocks'])"
