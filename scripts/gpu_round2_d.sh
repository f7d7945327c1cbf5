#!/bin/bash
# GPU-box script (round 2, call D, 1 GPU): full GPU suite without -x, batched-restart measurement, quick bench.
OUT=gpurun_out/${1:-r2d}
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/tests.log 2>&1; echo "exit $?" >> $OUT/tests.log
tail -12 $OUT/tests.log | cut -c1-300
timeout 400 python scripts/pair_bench.py 65536 10 > $OUT/pair_bench.log 2>&1; echo "exit $?" >> $OUT/pair_bench.log
tail -2 $OUT/pair_bench.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_quick.log 2>&1; echo "exit $?" >> $OUT/bench_quick.log
tail -2 $OUT/bench_quick.log | cut -c1-400
for W in readme3 dicty transform; do
  timeout 300 python bench.py --workload $W --steps 50 > $OUT/bench_$W.json 2> $OUT/bench_$W.err; echo "exit $?" >> $OUT/bench_$W.err
  tail -1 $OUT/bench_$W.json | cut -c1-700
done
