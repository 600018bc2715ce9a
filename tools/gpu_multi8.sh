#!/bin/bash
# 8-GPU box: multi-GPU parity test at world 2/4/8 + c2 and c5 bench lines at 8 ranks (in-run parity).  Usage: tools/gpu_multi8.sh <tag>
TAG=${1:-m8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/env.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -q -rs > $OUT/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -4 $OUT/pytest_multi.log
bash tools/gpu_multi2.sh $TAG 8
