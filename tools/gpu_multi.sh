#!/bin/bash
# Runs on an N-GPU box: multi-GPU parity test (world 2..N) + bench at 1..N ranks.  Usage: tools/gpu_multi.sh <tag> <N> [bench args]
TAG=${1:-m2}; N=${2:-2}; shift; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/env.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -q -rs > $OUT/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -8 $OUT/pytest_multi.log
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 "$@" > $OUT/bench_n1.json 2> $OUT/bench_n1.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 10 --warmup 3 "$@" > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err; fi
  echo "bench n=$n rc=$?"; cut -c1-600 $OUT/bench_n$n.json; tail -3 $OUT/bench_n$n.err
done
