#!/bin/bash
# quick check: aligner GPU parity tests + one timed C2 run + iteration profile.  Usage: tools/gpu_r2b.sh <tag> [pytest -k expr]
TAG=${1:-q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity_icp.py tests/test_gpu_golden.py -m gpu -q --maxfail=6 -k "${2:-not full_size}" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log
timeout 300 python tools/one_run.py 1000000 20 3 > $OUT/dbg.log 2>&1; grep -E "run 2" $OUT/dbg.log
timeout 300 python tools/iter_profile.py 1000000 8 > $OUT/iter_profile.txt 2>&1; cat $OUT/iter_profile.txt
