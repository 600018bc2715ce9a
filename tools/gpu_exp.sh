#!/bin/bash
TAG=${1:-exp}; OUT=$PWD/gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for r in 1 2; do timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> $OUT/err.txt | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"; done
timeout 600 python bench.py --config c5 --steps 5 --warmup 3 --no-cpu-baseline 2> $OUT/err.txt | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
