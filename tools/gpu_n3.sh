#!/bin/bash
TAG=${1:-n3}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_closure_batch.py -m gpu -x -q > $OUT/pytest_n3.log 2>&1; echo "pytest n3 rc=$?"; tail -30 $OUT/pytest_n3.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 300 python tools/closure_bench.py 16 20000 5000 2 > $OUT/closure_bench.txt 2>&1
timeout 300 python tools/closure_bench.py 16 100000 50000 3 >> $OUT/closure_bench.txt 2>&1
timeout 300 python tools/closure_bench.py 8 1000000 200000 3 >> $OUT/closure_bench.txt 2>&1
cat $OUT/closure_bench.txt
