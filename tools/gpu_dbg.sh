#!/bin/bash
# loop-kernel phase timing of one C2 run.  Usage: tools/gpu_dbg.sh <tag>
TAG=${1:-d}; OUT=gpurun_out/$TAG; mkdir -p $OUT
SRRG2B_LOOP_DEBUG=1 timeout 300 python tools/one_run.py 1000000 20 3 > $OUT/dbg.log 2>&1; tail -30 $OUT/dbg.log
