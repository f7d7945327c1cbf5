#!/bin/bash
# GPU-box script (round 2, call M, 1 GPU): what the driver runs at round end -- smoke, the GPU suite, the default bench -- plus
# the movielens completion line.
OUT=gpurun_out/${1:-r2m}
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "exit $?" >> $OUT/smoke.log
tail -2 $OUT/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/tests.log 2>&1; echo "exit $?" >> $OUT/tests.log
grep -E "passed|failed|FAILED|Error" $OUT/tests.log | head -12
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "exit $?" >> $OUT/bench_default.err
tail -1 $OUT/bench_default.json | cut -c1-400; tail -2 $OUT/bench_default.err
timeout 300 python bench.py --workload movielens --steps 50 > $OUT/bench_movielens.json 2> $OUT/bench_movielens.err; echo "exit $?" >> $OUT/bench_movielens.err
tail -1 $OUT/bench_movielens.json | cut -c1-700; tail -2 $OUT/bench_movielens.err
