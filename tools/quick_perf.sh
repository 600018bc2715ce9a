#!/bin/bash
# quick GPU check (via gpurun): parity tests, then per-iteration device times of the C2 run
TAG=${1:-q}
mkdir -p gpurun_out/$TAG
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/$TAG/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/$TAG/pytest_gpu.log
timeout 300 python tools/iter_profile.py 1000000 20 > gpurun_out/$TAG/iter_profile.txt 2>&1; tail -21 gpurun_out/$TAG/iter_profile.txt
if [ -n "${2:-}" ]; then
  SRRG2B_TILE=0 timeout 300 python tools/iter_profile.py 1000000 6 > gpurun_out/$TAG/iter_profile_notile.txt 2>&1; tail -6 gpurun_out/$TAG/iter_profile_notile.txt
fi
