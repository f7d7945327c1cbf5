#!/bin/bash
# GPU-box script: full GPU test-suite, ncu launch list + full capture of the fused kernels, default bench line.
OUT=gpurun_out/${1:-r01b}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/tests_gpu.log 2>&1; echo "exit $?" >> $OUT/tests_gpu.log
tail -15 $OUT/tests_gpu.log
if [ "$2" != "notests_only" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_n81920.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_fused_kernel -s 10 -c 2 -o $OUT/prof_fused_v3 \
    python bench.py --size 40960 --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_full_v3.log 2>&1
FZ_FUSED_VER=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_fused_t_kernel -s 10 -c 2 -o $OUT/prof_fused_v4 \
    python bench.py --size 40960 --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_full_v4.log 2>&1
ls -la $OUT
timeout 900 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "exit $?" >> $OUT/bench_n1.err
tail -1 $OUT/bench_n1.json | cut -c1-3000
fi
