#!/bin/bash
# GPU-box script: validate the v4 fused kernel (umma_fused_t.cuh) in isolation, then through the engine, then bench v3 vs v4.
# Every step runs under its own timeout and writes into gpurun_out/.
OUT=gpurun_out/v4
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
P=scikit-fusion_b200/csrc/dev/umma_probe
timeout 180 $P t 0 37888 > $OUT/probe_t_v0.log 2>&1; echo "exit $?" >> $OUT/probe_t_v0.log
if ! grep -q "(variant 0): 0 failing" $OUT/probe_t_v0.log; then
  timeout 120 $P t 1 0 > $OUT/probe_t_v1.log 2>&1; echo "exit $?" >> $OUT/probe_t_v1.log
  timeout 120 $P t 2 0 > $OUT/probe_t_v2_noB.log 2>&1; echo "exit $?" >> $OUT/probe_t_v2_noB.log
  timeout 120 $P t 4 0 > $OUT/probe_t_v4_noA.log 2>&1; echo "exit $?" >> $OUT/probe_t_v4_noA.log
  tail -30 $OUT/probe_t_v*.log
  exit 1
fi
tail -12 $OUT/probe_t_v0.log
FZ_FUSED_VER=4 timeout 600 python -m pytest tests -m gpu -x -q > $OUT/tests_v4.log 2>&1; echo "exit $?" >> $OUT/tests_v4.log
tail -5 $OUT/tests_v4.log
FZ_FUSED_VER=4 timeout 400 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_v4.log 2>&1; echo "exit $?" >> $OUT/bench_v4.log
tail -3 $OUT/bench_v4.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_v3.log 2>&1; echo "exit $?" >> $OUT/bench_v3.log
tail -3 $OUT/bench_v3.log
