#!/bin/bash
TAG=${1:-exp}; OUT=$PWD/gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_pgo.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py --config c4 --no-cpu-baseline > $OUT/bench_c4.json 2> $OUT/bench_c4.err; python -c "
import json; d=json.load(open('$OUT/bench_c4.json')); print(d['ms_per_step'], d['value'], d['result_check']['cg_iterations'], d['result_check']['solve_ms'])"
