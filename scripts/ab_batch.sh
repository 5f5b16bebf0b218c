#!/bin/bash
# A/B timing of the register kernel: library variants (scratch_libs/*.so) x chains per thread (MCMCB_K1_BATCH), C3 shape;
# DRAM bytes of one launch from ncu (local-memory spill traffic shows up here)
cp mcmcf90_b200/libmcmcb200.so /tmp/keep.so
for f in scratch_libs/*.so; do
  cp $f mcmcf90_b200/libmcmcb200.so
  for b in ${BATCHES:-1 2 4}; do
    echo "== $f batch=$b"; MCMCB_K1_BATCH=$b python scripts/quick_time.py 1048576 20 2>&1 | grep "N=" | tail -1 | cut -c1-140
    MCMCB_K1_BATCH=$b ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k1_step -s 1 -c 1 python scripts/prof_small.py 1048576 20 2>&1 | grep -E "dram__|gpu__time" 
  done
done
cp /tmp/keep.so mcmcf90_b200/libmcmcb200.so
