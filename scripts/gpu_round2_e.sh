#!/bin/bash
# GPU-box script (round 2, call E, 8 GPUs): the driver's N=8 line on the fixed n = 81 920 graph and BASELINE configs[3] literally.
OUT=gpurun_out/${1:-r2e}
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
FZ_GATE_LOG=1 timeout 500 $RUN bench.py --gpus 8 --steps 20 --warmup 3 > $OUT/bench_n8.log 2>&1; echo "exit $?" >> $OUT/bench_n8.log
tail -2 $OUT/bench_n8.log | cut -c1-2500
timeout 400 $RUN bench.py --gpus 8 --steps 20 --warmup 3 --size 100000 --no-e2e > $OUT/bench_n8_n100000.log 2>&1; echo "exit $?" >> $OUT/bench_n8_n100000.log
tail -2 $OUT/bench_n8_n100000.log | cut -c1-1200
