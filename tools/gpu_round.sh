#!/bin/bash
# Runs on the B200 box (via gpurun): smoke, GPU tests, bench, ncu launch list + full capture.
# Usage: tools/gpu_round.sh <tag> [skip_tests]
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/env.txt; nproc >> $OUT/env.txt; lscpu | grep 'Model name' >> $OUT/env.txt
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"
if [ "${2:-}" != "skip_tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nn_kernel|nn_far_kernel|linearize_kernel|icp_solve" -s 120 -c 4 -f -o $OUT/prof \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
