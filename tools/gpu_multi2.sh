#!/bin/bash
# N-GPU box: bench lines of c2 and c5 at N ranks with the in-run parity check.  Usage: tools/gpu_multi2.sh <tag> <N>
TAG=${1:-mm}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for cfg in c2 c5; do
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --config $cfg --steps 10 --warmup 3 > $OUT/bench_${cfg}_n$N.json 2> $OUT/bench_${cfg}_n$N.err; echo "bench $cfg n=$N rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench_${cfg}_n$N.json") if l.startswith("{")][-1])
print("$cfg", "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "parity", d["result_check"]["parity"])
PY
  tail -2 $OUT/bench_${cfg}_n$N.err
done
