#!/bin/bash
# round-2 GPU check: smoke + GPU tests + per-iteration profile + bench (no ncu).  Usage: tools/gpu_r2.sh <tag> [pytest-args]
TAG=${1:-q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/env.txt; nproc >> $OUT/env.txt
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 ${2:-} > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest_gpu.log
timeout 300 python tools/iter_profile.py > $OUT/iter_profile.txt 2>&1; cat $OUT/iter_profile.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-600 $OUT/bench.json; tail -3 $OUT/bench.err
