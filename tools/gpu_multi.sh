#!/bin/bash
# Runs on an N-GPU box: multi-GPU parity test + bench at N ranks.  Usage: tools/gpu_multi.sh <tag> <N>
TAG=${1:-m2}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/env.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $OUT/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -5 $OUT/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench rc=$?"; cat $OUT/bench_n$N.json; tail -5 $OUT/bench_n$N.err
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench1 rc=$?"; cat $OUT/bench_n1.json
