#!/bin/bash
# GPU-box script (round 2, call Z): x3 tests on the rebuilt library (tiled k x k chain, row-wise plane kernels), rank-128 timing,
# and the SM-reservation sweep of the persistent single-term kernel (FZ_RESERVE_SMS) on the default bench graph.
OUT=gpurun_out/${1:-r2z}
mkdir -p $OUT
timeout 420 python -m pytest tests/test_bf16x3_gpu.py tests/test_engine_parity.py tests/test_objective_gpu.py -q -m gpu --timeout 150 2>&1 | tail -30 > $OUT/x3_tests.log; echo "exit ${PIPESTATUS[0]}" >> $OUT/x3_tests.log
tail -6 $OUT/x3_tests.log
timeout 200 python scripts/x3_bench.py 16384 20 "rank128 cuda-core,rank128 planes,f32 planes" > $OUT/x3_rank128.jsonl 2> $OUT/x3_rank128.err; echo "exit $?" >> $OUT/x3_rank128.err
cut -c1-200 $OUT/x3_rank128.jsonl
for R in ${2:-0 4 8 12 16}; do
  FZ_RESERVE_SMS=$R timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_reserve$R.log 2>&1; echo "exit $?" >> $OUT/bench_reserve$R.log
  python - $OUT/bench_reserve$R.log $R <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print("reserve", sys.argv[2], "it/s", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "share", d["roofline"]["kernel_share_of_step"], "mhz", d["clocks"]["sm_mhz"], "W", d["clocks"].get("power_w"))
PY
done
