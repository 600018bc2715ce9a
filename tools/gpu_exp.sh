#!/bin/bash
# experiment: threshold between warp-per-query and thread-per-query handling of a coherence work list
TAG=${1:-exp}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for f in 6 7 8 10; do
  SRRG2B_SMALL_SHIFT=$f timeout 300 python tools/iter_profile.py 1000000 6 > $OUT/iter_profile_$f.txt 2>&1; echo "shift=$f"; cat $OUT/iter_profile_$f.txt
done
