#!/bin/bash
# GPU-box script (round 2, call Y): whole GPU suite on the rebuilt library, float32-at-scale timings (bf16 planes), launch list of
# the planes path under ncu.
OUT=gpurun_out/${1:-r2y}
mkdir -p $OUT
timeout 600 python -m pytest tests -q -m gpu --timeout 200 2>&1 | tail -40 > $OUT/gpu_tests.log; echo "exit ${PIPESTATUS[0]}" >> $OUT/gpu_tests.log
tail -12 $OUT/gpu_tests.log
timeout 300 python scripts/x3_bench.py 16384 20 > $OUT/x3_bench.jsonl 2> $OUT/x3_bench.err; echo "exit $?" >> $OUT/x3_bench.err
python - <<'PY'
import json
for l in open("gpurun_out/r2y/x3_bench.jsonl"):
    d = json.loads(l); print(d["workload"], d["it_per_s"], d["launches_per_it"], d.get("relFro_G_vs_cuda_core"))
PY
tail -3 $OUT/x3_bench.err
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_x3_planes_n16384.csv \
  python scripts/x3_bench.py 16384 2 "f32 planes,rank128 planes" > $OUT/ncu_x3.log 2>&1; echo "exit $?" >> $OUT/ncu_x3.log
tail -2 $OUT/ncu_x3.log | cut -c1-200
