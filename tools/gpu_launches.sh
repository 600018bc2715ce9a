#!/bin/bash
# in-situ per-kernel durations (single-pass ncu, no cache flush, no clock control).  Usage: tools/gpu_launches.sh <tag>
TAG=${1:-l}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 700 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 1 --warmup 2 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python tools/launch_table.py $OUT/launches.csv | tee $OUT/launch_table.txt
