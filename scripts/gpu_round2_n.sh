#!/bin/bash
# GPU-box script (round 2, call N): the driver's bench line at N GPUs (n = 81 920, e2e leg included).
N=${2:-4}
OUT=gpurun_out/${1:-r2n}
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581"
FZ_GATE_LOG=1 timeout 500 $RUN bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n$N.log 2>&1; echo "exit $?" >> $OUT/bench_n$N.log
grep "fz peer" $OUT/bench_n$N.log | sort | uniq -c | head -2
grep '^{' $OUT/bench_n$N.log | cut -c1-250
tail -1 $OUT/bench_n$N.log
