#!/bin/bash
TAG=${1:-exp}; OUT=$PWD/gpurun_out/$TAG; mkdir -p $OUT
show() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"; }
echo "== c5 old"; (cd _old && timeout 600 python bench.py --config c5 --steps 5 --warmup 3 --no-cpu-baseline 2> $OUT/err_old.txt | show)
echo "== c5 new"; timeout 600 python bench.py --config c5 --steps 5 --warmup 3 --no-cpu-baseline 2> $OUT/err.txt | show
echo "== c5 new, far ctas 8"; SRRG2B_FAR_SOLE_CTAS=8 timeout 600 python bench.py --config c5 --steps 5 --warmup 3 --no-cpu-baseline 2> $OUT/err.txt | show
echo "== c2 old"; (cd _old && timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> $OUT/err_old.txt | show)
echo "== c2 new"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> $OUT/err.txt | show
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
