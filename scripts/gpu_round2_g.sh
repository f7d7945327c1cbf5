#!/bin/bash
# GPU-box script (round 2, call G, N GPUs): multi-GPU tests (when N == 2) and the A/B of the dynamic schedule tail under NCCL.
N=${2:-4}
OUT=gpurun_out/${1:-r2g}
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561"
for D in 0; do
  FZ_DYN_SCHED=$D timeout 500 $RUN bench.py --gpus $N --steps 20 --warmup 3 --no-e2e > $OUT/bench_n${N}_dyn$D.log 2>&1; echo "exit $?" >> $OUT/bench_n${N}_dyn$D.log
  grep '^{' $OUT/bench_n${N}_dyn$D.log | cut -c1-200
done
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_objective_gpu.py -m gpu -q > $OUT/tests_multi.log 2>&1; echo "exit $?" >> $OUT/tests_multi.log
  tail -5 $OUT/tests_multi.log
fi
