#!/bin/bash
# GPU-box script: smoke(), full GPU suite (default kernels and FZ_FUSED_VER=4), quick bench with v4.
OUT=gpurun_out/${1:-final}
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "exit $?" >> $OUT/smoke.log; tail -2 $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/tests_gpu.log 2>&1; echo "exit $?" >> $OUT/tests_gpu.log; tail -4 $OUT/tests_gpu.log
FZ_FUSED_VER=4 timeout 900 python -m pytest tests -m gpu -q > $OUT/tests_gpu_v4.log 2>&1; echo "exit $?" >> $OUT/tests_gpu_v4.log; tail -4 $OUT/tests_gpu_v4.log
FZ_FUSED_VER=4 timeout 400 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_v4.log 2>&1; tail -1 $OUT/bench_v4.log | cut -c1-400
