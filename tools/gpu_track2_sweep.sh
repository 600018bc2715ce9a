#!/bin/bash
# iteration profile of the C2 run under different bound-certification schedules.  Usage: tools/gpu_track2_sweep.sh <tag>
TAG=${1:-tr}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for cfg in "default" "SRRG2B_TRACK2_FRAC=4" "SRRG2B_TRACK2=1" "SRRG2B_PRE_ITERS=2 SRRG2B_TRACK2_FRAC=4" "SRRG2B_PRE_ITERS=1 SRRG2B_TRACK2=1" "SRRG2B_PRE_ITERS=4"; do
  echo "== $cfg" | tee -a $OUT/sweep.txt
  if [ "$cfg" = "default" ]; then timeout 300 python tools/iter_profile.py 1000000 8 >> $OUT/sweep.txt 2>&1
  else env $cfg timeout 300 python tools/iter_profile.py 1000000 8 >> $OUT/sweep.txt 2>&1; fi
done
cat $OUT/sweep.txt
