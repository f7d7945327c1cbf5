#!/bin/bash
# GPU-box script (round 2, call X): exact bf16 planes (storage bfloat16x3) -- parity tests, then float32-at-scale timings.
OUT=gpurun_out/${1:-r2x}
mkdir -p $OUT
timeout 420 python -m pytest tests/test_bf16x3_gpu.py -q -m gpu --timeout 150 2>&1 | tail -60 > $OUT/x3_tests.log; echo "exit ${PIPESTATUS[0]}" >> $OUT/x3_tests.log
tail -25 $OUT/x3_tests.log
if [ "${2:-bench}" = "bench" ]; then
  timeout 300 python scripts/x3_bench.py ${3:-16384} 20 > $OUT/x3_bench.jsonl 2> $OUT/x3_bench.err; echo "exit $?" >> $OUT/x3_bench.err
  cat $OUT/x3_bench.jsonl | cut -c1-330
  tail -3 $OUT/x3_bench.err
fi
