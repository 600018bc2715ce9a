#!/bin/bash
# ncu --set full captures of selected kernels of one C2 run.  Usage: tools/gpu_ncu.sh <tag> <kernel-regex> <skip> <count>
TAG=${1:-n}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s ${3:-0} -c ${4:-2} -f -o $OUT/prof \
  python tools/one_run.py 1000000 20 2 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"; tail -5 $OUT/ncu.log
ls -la $OUT
