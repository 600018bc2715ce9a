#!/bin/bash
# experiment: flat row chunks in the thread-per-query searches (SRRG2B_NN_FLAT bit 0: phase 1, bit 1: far / full lists)
TAG=${1:-flat}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity_icp.py tests/test_gpu_golden.py -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
for f in 0 3; do
  SRRG2B_NN_FLAT=$f timeout 300 python tools/iter_profile.py 1000000 5 > $OUT/iter_profile_flat$f.txt 2>&1; echo "flat=$f"; cat $OUT/iter_profile_flat$f.txt
  SRRG2B_NN_FLAT=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_flat$f.json 2> $OUT/bench_flat$f.err; echo "bench rc=$?"; cut -c1-200 $OUT/bench_flat$f.json
done
