#!/bin/bash
# quick GPU check: smoke + ICP parity tests + iteration profile + bench (no ncu).  Usage: tools/gpu_quick.sh <tag>
TAG=${1:-q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 300 python tools/iter_profile.py > $OUT/iter_profile.txt 2>&1; cat $OUT/iter_profile.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-400 $OUT/bench.json
