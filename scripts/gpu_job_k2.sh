#!/bin/bash
# K2/K3 iteration job: parity tests of the warp-per-chain kernels, timings, per-launch metrics
python -m pytest tests/test_k2_parity.py tests/test_k3_scam_parity.py tests/test_pool_diag.py -q -x 2>&1 | tail -4
python scripts/time_k2.py ${1:-all} 2>&1 | grep -v "^$" | cut -c1-200 | awk "NR%3==0"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,dram__bytes_read.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_k2.csv python scripts/time_k2.py ${1:-all} > /dev/null 2>&1
python scripts/launch_table.py gpurun_out/launches_k2.csv | awk '$3 > 1000000'
