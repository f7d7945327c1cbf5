#!/bin/bash
# GPU-box script: correctness + burst + sustained study of the single-term fused kernel (v5) next to v3.
OUT=gpurun_out/${1:-v5probe}
mkdir -p $OUT
P=scikit-fusion_b200/csrc/dev/umma_probe
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $OUT/smi.csv &
SMI=$!
timeout 400 $P 1 ${2:-37888} ${3:-4} ${4:-2} > $OUT/probe_v5.log 2>&1; echo "exit $?" >> $OUT/probe_v5.log
kill $SMI
tail -30 $OUT/probe_v5.log
timeout 300 python scripts/v5_engine_check.py > $OUT/engine_check.log 2>&1; echo "exit $?" >> $OUT/engine_check.log
cat $OUT/engine_check.log | tail -20
for M in centred1 2; do
  timeout 400 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --split-terms $M > $OUT/bench_$M.log 2>&1; echo "exit $?" >> $OUT/bench_$M.log
  tail -2 $OUT/bench_$M.log | cut -c1-1500
done
