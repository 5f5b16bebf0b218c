#!/bin/bash
# C3 ncu capture at the final source hash, 20 of the 100 iterations of a bench step (bench.py scales by chain-steps)
mkdir -p gpurun_out
timeout 105 python scripts/ncu_profile.py c3 --iters 20 > gpurun_out/r02_ncu_c3.log 2>&1 || tail -n 5 gpurun_out/r02_ncu_c3.log
ls -la gpurun_out/r02_ncu_c3.json
