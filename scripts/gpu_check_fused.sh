#!/bin/bash
# GPU-box script: probe the v4 fused kernel, run the GPU test-suite with both fused kernels, bench both (short).
OUT=gpurun_out/${1:-check}
mkdir -p $OUT
P=scikit-fusion_b200/csrc/dev/umma_probe
timeout 150 $P t 0 37888 > $OUT/probe_v4.log 2>&1; echo "exit $?" >> $OUT/probe_v4.log
tail -9 $OUT/probe_v4.log
for V in 4 3; do
  FZ_FUSED_VER=$V timeout 600 python -m pytest tests -m gpu -x -q > $OUT/tests_v$V.log 2>&1; echo "exit $?" >> $OUT/tests_v$V.log
  tail -3 $OUT/tests_v$V.log
  FZ_FUSED_VER=$V timeout 400 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_v$V.log 2>&1; echo "exit $?" >> $OUT/bench_v$V.log
  tail -2 $OUT/bench_v$V.log | cut -c1-2000
done
