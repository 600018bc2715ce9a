#!/bin/bash
# One 1-GPU evidence pass of a round: smoke, GPU tests, the bench lines of every configuration + the reference arm,
# the ncu launch list of the bench command and one `ncu --set full` capture of the dominant kernel.
# Usage: tools/gpu_round.sh <tag>
TAG=${1:-round}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $OUT/env.txt; nproc >> $OUT/env.txt
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu.log
for cfg in c2 c5 c3 c4; do
  extra=""; [ $cfg = c3 ] && extra="--steps 2"
  timeout 900 python bench.py --config $cfg $extra > $OUT/bench_$cfg.json 2> $OUT/bench_$cfg.err; echo "bench $cfg rc=$?"; cut -c1-200 $OUT/bench_$cfg.json
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference_c2.json 2> $OUT/bench_reference_c2.err; echo "reference rc=$?"; cut -c1-200 $OUT/bench_reference_c2.json
timeout 300 python tools/closure_bench.py 16 20000 5000 2 > $OUT/closure_bench.txt 2>&1; timeout 300 python tools/closure_bench.py 16 100000 50000 3 >> $OUT/closure_bench.txt 2>&1; cat $OUT/closure_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 900 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 1 --warmup 2 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python tools/launch_table.py $OUT/launches.csv > $OUT/launch_table.txt; head -8 $OUT/launch_table.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:check_tiles_kernel -s 30 -c 2 -f -o $OUT/check_tiles \
  python tools/one_run.py 1000000 20 2 > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_summary.py $OUT/check_tiles.ncu-rep > $OUT/ncu_check_tiles_summary.txt 2>&1; head -12 $OUT/ncu_check_tiles_summary.txt
rm -f $OUT/check_tiles.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nn_kernel|nn_far_kernel" -s 0 -c 6 -f -o $OUT/nn \
  python tools/one_run.py 1000000 4 1 > $OUT/ncu_nn.log 2>&1; echo "ncu nn rc=$?"
for k in 0 1 2; do python tools/ncu_lines2.py $OUT/nn.ncu-rep 14 $k; done > $OUT/ncu_nn_lines.txt 2>&1
python tools/ncu_summary.py $OUT/nn.ncu-rep > $OUT/ncu_nn_summary.txt 2>&1; head -12 $OUT/ncu_nn_summary.txt
rm -f $OUT/nn.ncu-rep
