#!/bin/bash
# GPU-box script (round 2, call T): phase timeline of the sharded iteration (FZ_TIMELINE=1) at N GPUs.
N=${2:-4}
OUT=gpurun_out/${1:-r2t}
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591"
FZ_TIMELINE=1 timeout 400 $RUN bench.py --gpus $N --steps 20 --warmup 3 --no-e2e > $OUT/bench_n$N.log 2>&1; echo "exit $?" >> $OUT/bench_n$N.log
grep "fz timeline" $OUT/bench_n$N.log | head -3
grep '^{' $OUT/bench_n$N.log | cut -c1-200
tail -1 $OUT/bench_n$N.log
