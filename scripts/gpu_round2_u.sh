#!/bin/bash
# GPU-box script (round 2, call U, 2 GPUs): the multi-GPU tests on the final library, then a short 2-GPU bench (peer exchange at
# full size; its `check` must equal the 1-GPU value).
OUT=gpurun_out/${1:-r2u}
mkdir -p $OUT
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 200 2>&1 | tail -15 > $OUT/tests_multi.log; echo "exit ${PIPESTATUS[0]}" >> $OUT/tests_multi.log
tail -5 $OUT/tests_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 13 --no-e2e --no-cpu > $OUT/bench_n2.log 2>&1; echo "exit $?" >> $OUT/bench_n2.log
grep '^{' $OUT/bench_n2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=2', d['value'], 'it/s', d['ms_per_step'], 'ms; frac', d['roofline']['frac'], 'check', d['check'])"
tail -1 $OUT/bench_n2.log
