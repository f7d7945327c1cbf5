#!/bin/bash
# GPU-box script (round 2, call F, 1 GPU): probe of the hybrid schedule, full GPU suite, quick benches with / without the
# dynamic tail and the pre-split.
OUT=gpurun_out/${1:-r2f}
mkdir -p $OUT
P=scikit-fusion_b200/csrc/dev/umma_probe
timeout 300 $P 1 37888 3 148 > $OUT/probe_v5.log 2>&1; echo "exit $?" >> $OUT/probe_v5.log
grep -E "bench|correctness|FAIL" $OUT/probe_v5.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/tests.log 2>&1; echo "exit $?" >> $OUT/tests.log
grep -E "passed|failed|FAILED" $OUT/tests.log | head -12
timeout 400 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_quick.log 2>&1; echo "exit $?" >> $OUT/bench_quick.log
tail -2 $OUT/bench_quick.log | cut -c1-300
FZ_NO_DYN_SCHED=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_quick_static.log 2>&1; echo "exit $?" >> $OUT/bench_quick_static.log
tail -2 $OUT/bench_quick_static.log | cut -c1-300
timeout 300 python bench.py --workload transform --steps 5 > $OUT/bench_transform.json 2> $OUT/bench_transform.err; echo "exit $?" >> $OUT/bench_transform.err
tail -1 $OUT/bench_transform.json | cut -c1-900
