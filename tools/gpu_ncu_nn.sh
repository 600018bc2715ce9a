#!/bin/bash
# ncu --set full of the search kernels of the first run's iterations 1-3 (cold), flat off / on
TAG=${1:-nn}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for f in 0 3; do
SRRG2B_NN_FLAT=$f timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nn_kernel|nn_far_kernel" -s 0 -c 6 -f -o $OUT/prof_flat$f \
  python tools/one_run.py 1000000 4 1 > $OUT/ncu_flat$f.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/ncu_flat$f.log
done
ls -la $OUT
