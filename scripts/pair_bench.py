"""GPU-box measurement: restarts batched two per pass over the relations
(fz_pair_iterate) against the same two restarts run one after the other, on the bench graph.  Prints one JSON line.
    python scripts/pair_bench.py [n_per_type] [iterations]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-fusion_b200"))
import torch  # noqa: E402
from skfusion import _capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
T, K = 5, 64
pairs = [(i, j) for i in range(T) for j in range(T) if i < j]
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream(dev).cuda_stream
R = {}
for (i, j) in pairs:
    t = torch.empty((n, n), dtype=torch.bfloat16, device=dev)
    _capi.fill_uniform(t, 1000 + 10 * i + j, stream=st)
    R[i, j] = t
rs = np.random.RandomState(0)


def handle(first=None):
    eng = _capi.Engine(0, "float32")
    eng.set_split_terms("centred1")
    tid = [eng.add_type(n, K) for _ in range(T)]
    for r, (i, j) in enumerate(pairs):
        if first is None:
            eng.add_relation(tid[i], tid[j], R[i, j], storage="bfloat16", borrow=True)
        else:
            eng.add_relation_borrowed(tid[i], tid[j], *first.relation_device_ptr(r))
    for t in tid:
        eng.set_factor(t, rs.rand(n, K).astype(np.float32))
    eng.finalize()
    return eng


def timed(fn):
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1)


a = handle()
b = handle(a)
a.iterate(_capi.FZ_DFMF, 3, st)
b.iterate(_capi.FZ_DFMF, 3, st)
seq_ms = timed(lambda: (a.iterate(_capi.FZ_DFMF, iters, st), b.iterate(_capi.FZ_DFMF, iters, st)))
a.pair_iterate(b, 3, st)
pair_ms = timed(lambda: a.pair_iterate(b, iters, st))
stats = a.operand_stats()
print(json.dumps({"n_per_type": n, "iterations": iters, "relation_bytes": 10 * n * n * 2,
                  "sequential_ms_per_iteration_per_run": seq_ms / (2 * iters), "paired_ms_per_iteration_per_run": pair_ms / (2 * iters),
                  "restart_throughput_gain": seq_ms / pair_ms, "run_iterations_per_s_sequential": 2000.0 * iters / seq_ms,
                  "run_iterations_per_s_paired": 2000.0 * iters / pair_ms, "paired_iterations": stats["paired"]}))
b.close()
a.close()
