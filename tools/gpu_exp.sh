#!/bin/bash
# experiment: rounds of in-place failure searches inside check_tiles_kernel (library variants built with -DS2B_FAIL_ROUNDS)
TAG=${1:-exp}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -m pytest tests/test_cpp_host_mirror.py -m gpu -x -q > $OUT/pytest_cpp.log 2>&1; echo "pytest cpp rc=$?"; tail -15 $OUT/pytest_cpp.log
for v in "" _fr1 _fr0; do
  export SRRG2B_LIB=$PWD/srrg2_slam_interfaces_b200/libsrrg2b$v.so
  timeout 300 python tools/iter_profile.py 1000000 8 > $OUT/iter_profile$v.txt 2>&1; echo "variant=$v"; tail -4 $OUT/iter_profile$v.txt
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench$v.json 2> $OUT/bench$v.err; echo "bench rc=$?"; cut -c1-200 $OUT/bench$v.json
done
