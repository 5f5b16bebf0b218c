#!/bin/bash
# --set full of ONE wave of the theta-in-shared-memory SCAM kernel (18 944 chains = 148 CTAs, 3 sweeps), 4 and 2 lanes
mkdir -p gpurun_out
for L in 4 2; do
  MCMCB_K5S_LANES=$L timeout 400 ncu --set full --clock-control none --import-source on -k regex:k5s_scam_step_kernel -s 1 -c 1 -o gpurun_out/r02_k5s_full_L$L -f python scripts/ncu_profile.py c5 --child --chains 18944 --iters 3 > gpurun_out/r02_k5s_full_L$L.log 2>&1
  ncu -i gpurun_out/r02_k5s_full_L$L.ncu-rep --page details > gpurun_out/r02_ncu_k5s_full_L${L}_details.txt 2>&1
  ncu -i gpurun_out/r02_k5s_full_L$L.ncu-rep --page source --csv > gpurun_out/r02_ncu_k5s_full_L${L}_source.csv 2>&1
  rm -f gpurun_out/r02_k5s_full_L$L.ncu-rep
  grep -E "Duration|Executed Ipc Active|Issue Slots Busy|No Eligible|Eligible Warps|Achieved Occupancy|FP64|Registers Per|Dynamic Shared|L1/TEX Hit|Bank conflicts|Stall" gpurun_out/r02_ncu_k5s_full_L${L}_details.txt | head -30
done
ls -la gpurun_out | tail -8
