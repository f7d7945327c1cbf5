#!/bin/bash
# GPU-box script (round 2, call V): final 1-GPU validation of the shipped library -- smoke(), the whole GPU suite, the two-pass
# (umma_skinny) workloads of scripts/x3_bench.py after the elected-lane issue change.
OUT=gpurun_out/${1:-r2v}
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "exit $?" >> $OUT/smoke.log; tail -2 $OUT/smoke.log
timeout 700 python -m pytest tests -q -m gpu --timeout 200 2>&1 | tail -40 > $OUT/gpu_tests.log; echo "exit ${PIPESTATUS[0]}" >> $OUT/gpu_tests.log
tail -8 $OUT/gpu_tests.log
timeout 300 python scripts/x3_bench.py 16384 20 "rank128 cuda-core,rank128 planes,completion cuda-core,completion planes,f32 cuda-core,f32 planes" > $OUT/x3_bench.jsonl 2> $OUT/x3_bench.err; echo "exit $?" >> $OUT/x3_bench.err
python - <<'PY'
import json
for l in open("gpurun_out/r2v/x3_bench.jsonl"):
    d = json.loads(l); print(d["workload"], d["it_per_s"], d["launches_per_it"], d.get("relFro_G_vs_cuda_core"))
PY
tail -2 $OUT/x3_bench.err
