#!/bin/bash
# GPU-box script (round 2, call A): persistent single-term kernel probe, engine check with the gate log, GPU suite, benches, launch list.
OUT=gpurun_out/${1:-r2a}
mkdir -p $OUT
P=scikit-fusion_b200/csrc/dev/umma_probe
timeout 300 $P 1 37888 3 148 > $OUT/probe_v5.log 2>&1; echo "exit $?" >> $OUT/probe_v5.log
grep -E "bench|correctness|FAIL" $OUT/probe_v5.log
FZ_GATE_LOG=1 timeout 300 python scripts/v5_engine_check.py > $OUT/engine_check.log 2>&1; echo "exit $?" >> $OUT/engine_check.log
tail -25 $OUT/engine_check.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/tests.log 2>&1; echo "exit $?" >> $OUT/tests.log
tail -5 $OUT/tests.log
for M in auto centred1; do
  FZ_GATE_LOG=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --split-terms $M > $OUT/bench_$M.log 2>&1; echo "exit $?" >> $OUT/bench_$M.log
  tail -2 $OUT/bench_$M.log | cut -c1-1200
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 300 --csv --log-file $OUT/launches_auto.csv python bench.py --steps 3 --warmup 4 --no-e2e --no-cpu --split-terms auto > $OUT/ncu_bench.log 2>&1; echo "exit $?" >> $OUT/ncu_bench.log
tail -2 $OUT/ncu_bench.log | cut -c1-300
