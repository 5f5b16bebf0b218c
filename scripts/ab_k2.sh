#!/bin/bash
# A/B timing of library variants (scratch_libs/*.so) on the large-npar shapes
cp mcmcf90_b200/libmcmcb200.so /tmp/keep.so
for f in /tmp/keep.so scratch_libs/*.so; do
  cp $f mcmcf90_b200/libmcmcb200.so
  for cfg in ${CFGS:-c2 c4 c5}; do echo "== $f $cfg"; timeout 300 python scripts/time_k2.py $cfg 2>&1 | grep "steps=" | tail -1 | cut -c1-110; done
done
cp /tmp/keep.so mcmcf90_b200/libmcmcb200.so
