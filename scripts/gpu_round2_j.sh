#!/bin/bash
# GPU-box script (round 2, call J, N GPUs): the peer-memory reduce-scatter -- multi-GPU tests, then bench with it and with NCCL's.
N=${2:-2}
OUT=gpurun_out/${1:-r2j}
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571"
if [ "$N" = "2" ]; then
  FZ_GATE_LOG=1 timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > $OUT/tests_multi.log 2>&1; echo "exit $?" >> $OUT/tests_multi.log
  grep -E "passed|failed|FAILED|Error|fz peer" $OUT/tests_multi.log | sort | uniq -c | head -12
fi
for P in 1 0; do
  FZ_GATE_LOG=1 FZ_PEER_RS=$P timeout 500 $RUN bench.py --gpus $N --steps 20 --warmup 3 --no-e2e > $OUT/bench_n${N}_peer$P.log 2>&1; echo "exit $?" >> $OUT/bench_n${N}_peer$P.log
  grep "fz peer" $OUT/bench_n${N}_peer$P.log | sort | uniq -c | head -3
  grep '^{' $OUT/bench_n${N}_peer$P.log | cut -c1-200
  tail -2 $OUT/bench_n${N}_peer$P.log | cut -c1-300
done
