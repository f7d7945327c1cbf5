#!/bin/bash
# GPU-box script (round 2, call H, 1 GPU): full GPU suite, quick bench, launch list (DMMA reductions, trace objective, pre-split).
OUT=gpurun_out/${1:-r2h}
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/tests.log 2>&1; echo "exit $?" >> $OUT/tests.log
grep -E "passed|failed|FAILED|Error" $OUT/tests.log | head -12
timeout 400 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_quick.log 2>&1; echo "exit $?" >> $OUT/bench_quick.log
tail -2 $OUT/bench_quick.log | cut -c1-300
FZ_NO_DMMA=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_quick_nodmma.log 2>&1; echo "exit $?" >> $OUT/bench_quick_nodmma.log
tail -2 $OUT/bench_quick_nodmma.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 300 --csv --log-file $OUT/launches_auto.csv python bench.py --steps 3 --warmup 5 --no-e2e --no-cpu > $OUT/ncu_bench.log 2>&1; echo "exit $?" >> $OUT/ncu_bench.log
tail -1 $OUT/ncu_bench.log | cut -c1-100
