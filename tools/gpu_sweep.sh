#!/bin/bash
# iteration profile of the C2 run under a list of environment settings.  Usage: tools/gpu_sweep.sh <tag> "<env1>" "<env2>" ...
TAG=${1:-sw}; OUT=gpurun_out/$TAG; mkdir -p $OUT; shift
for cfg in "$@"; do
  echo "== $cfg" | tee -a $OUT/sweep.txt
  if [ "$cfg" = "default" ]; then timeout 300 python tools/iter_profile.py 1000000 ${ITERS:-8} >> $OUT/sweep.txt 2>&1
  else env $cfg timeout 300 python tools/iter_profile.py 1000000 ${ITERS:-8} >> $OUT/sweep.txt 2>&1; fi
done
cat $OUT/sweep.txt
