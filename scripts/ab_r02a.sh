#!/bin/bash
# round 2, experiment A: direct exp table (8-byte / split), 256-thread CTAs, against the masked-table baseline
mkdir -p gpurun_out
out=gpurun_out/r02_abA.txt; : > $out
cp mcmcf90_b200/libmcmcb200.so /tmp/keep.so
run() { # label lib env...
  label=$1; lib=$2; shift 2
  cp $lib mcmcf90_b200/libmcmcb200.so
  echo "== $label" >> $out
  env "$@" python scripts/quick_time.py 1048576 100 2>&1 | grep "N=" | tail -n 2 | cut -c1-150 >> $out
  env "$@" ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k1_step -s 1 -c 1 python scripts/prof_small.py 1048576 100 2>&1 | grep -E "dram__|gpu__time|l1tex|smsp__|sm__" >> $out
}
run base /tmp/keep.so MCMCB_EXP_DIRECT=0
run direct8 /tmp/keep.so MCMCB_EXP_DIRECT=1
run split scratch_libs/split.so MCMCB_EXP_DIRECT=1
run t256_direct8 scratch_libs/t256.so MCMCB_EXP_DIRECT=1
run t256_base scratch_libs/t256.so MCMCB_EXP_DIRECT=0
cp /tmp/keep.so mcmcf90_b200/libmcmcb200.so
cat $out
