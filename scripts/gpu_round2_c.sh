#!/bin/bash
# GPU-box script (round 2, call C, 1 GPU): full GPU suite, the driver's default bench line, the reference arm, ncu captures.
OUT=gpurun_out/${1:-r2c}
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/tests.log 2>&1; echo "exit $?" >> $OUT/tests.log
tail -6 $OUT/tests.log
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "exit $?" >> $OUT/bench_default.err
tail -1 $OUT/bench_default.json | cut -c1-600; tail -3 $OUT/bench_default.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "exit $?" >> $OUT/bench_reference.err
tail -1 $OUT/bench_reference.json | cut -c1-1500
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_fused1 -s 12 -c 2 -o $OUT/ncu_fused1 python bench.py --size 40960 --steps 2 --warmup 2 --no-e2e --no-cpu > $OUT/ncu_full.log 2>&1; echo "exit $?" >> $OUT/ncu_full.log
tail -2 $OUT/ncu_full.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 300 --csv --log-file $OUT/launches_auto.csv python bench.py --steps 3 --warmup 5 --no-e2e --no-cpu > $OUT/ncu_bench.log 2>&1; echo "exit $?" >> $OUT/ncu_bench.log
tail -1 $OUT/ncu_bench.log | cut -c1-200
