#!/bin/bash
# A/B timing of library variants built with different -D flags (scratch_libs/*.so), C3 shape
for f in default scratch_libs/*.so; do
  if [ "$f" != default ]; then cp mcmcf90_b200/libmcmcb200.so /tmp/keep.so; cp $f mcmcf90_b200/libmcmcb200.so; fi
  echo "== $f"; python scripts/quick_time.py 1048576 20 2>&1 | grep "N=" | tail -1 | cut -c1-150
  if [ "$f" != default ]; then cp /tmp/keep.so mcmcf90_b200/libmcmcb200.so; fi
done
