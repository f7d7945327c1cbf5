#!/bin/bash
# GPU-box script (round 2, call B, 2 GPUs): multi-GPU tests (torchrun + one-process group), profile products, 2-GPU bench.
OUT=gpurun_out/${1:-r2b}
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_profile_products_gpu.py -m gpu -x -q > $OUT/tests_multi.log 2>&1; echo "exit $?" >> $OUT/tests_multi.log
tail -15 $OUT/tests_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --split-terms auto --size 49152 > $OUT/bench_n2.log 2>&1; echo "exit $?" >> $OUT/bench_n2.log
tail -3 $OUT/bench_n2.log | cut -c1-3000
