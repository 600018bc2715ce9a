#!/bin/bash
# 1-GPU bench lines of every configuration.  Usage: tools/gpu_bench_all.sh <tag>
TAG=${1:-ba}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for cfg in c2 c5 c3 c4; do
  extra=""; [ $cfg = c4 ] && extra="--steps 6"; [ $cfg = c3 ] && extra="--steps 2"
  timeout 900 python bench.py --config $cfg $extra > $OUT/bench_$cfg.json 2> $OUT/bench_$cfg.err; echo "bench $cfg rc=$?"; cut -c1-300 $OUT/bench_$cfg.json; tail -3 $OUT/bench_$cfg.err
done
