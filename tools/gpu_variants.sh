#!/bin/bash
# runs tools/one_run.py with the loop debug for every library variant gpurun_variants_*.so
OUT=gpurun_out/${1:-v}; mkdir -p $OUT
cp srrg2_slam_interfaces_b200/libsrrg2b.so /tmp/lib_orig.so
for f in gpurun_variants_*.so; do
  cp $f srrg2_slam_interfaces_b200/libsrrg2b.so
  echo "=== $f"
  SRRG2B_LOOP_DEBUG=1 timeout 300 python tools/one_run.py 1000000 20 3 2>&1 | grep -E "run 2|it 1[0-2]"
done
cp /tmp/lib_orig.so srrg2_slam_interfaces_b200/libsrrg2b.so
