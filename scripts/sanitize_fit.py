"""Small fits through every tensor-core path, meant to run under `compute-sanitizer --tool memcheck` (race /
memory checking of the unit kernels).  No oracle, no torch: numpy inputs through the seam functions; prints what ran."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-fusion_b200"))
from skfusion.fusion import solver     # noqa: E402

warnings.simplefilter("ignore")
rs = np.random.RandomState(0)
types = ["a", "b", "c"]
sizes = {"a": 392, "b": 264, "c": 136}
R = {(i, j): [rs.rand(sizes[i], sizes[j]).astype(np.float32).astype(np.float64)] for i, j in (("a", "b"), ("a", "c"), ("b", "c"))}
th = np.where(rs.rand(392, 392) < 0.02, -0.01, 0.0)
Th = {("a", "a"): [(th + th.T) / 2]}
its = int(sys.argv[1]) if len(sys.argv) > 1 else 2
runs = [("bf16 fused, centred two-term", dict(storage="bfloat16", split_terms="auto"), {"a": 64, "b": 64, "c": 64}, {}, None),
        ("bf16 fused, single-term", dict(storage="bfloat16", split_terms="centred1"), {"a": 64, "b": 40, "c": 24}, {}, None),
        ("exact planes, constraint, rank 96 (two-pass, N = 256)", dict(storage="bfloat16x3"), {"a": 96, "b": 40, "c": 24}, Th, None),
        ("exact planes, fused, constraint", dict(storage="bfloat16x3", split_terms="auto"), {"a": 48, "b": 40, "c": 24}, Th, None),
        ("exact planes, completion", dict(storage="bfloat16x3"), {"a": 24, "b": 32, "c": 8}, {},
         {("a", "b"): [rs.rand(392, 264) < 0.3], ("a", "c"): [None], ("b", "c"): [None]})]
for label, kw, ranks, theta, M in runs:
    if M is None:
        G, S = solver.dfmf(R, theta, types, ranks, max_iter=its, init_type="random", random_state=np.random.RandomState(1), dtype="float32", **kw)
    else:
        G, S = solver.dfmc(R, M, theta, types, ranks, max_iter=its, init_type="random", random_state=np.random.RandomState(1), dtype="float32", **kw)
    ok = all(np.isfinite(v).all() for v in G.values())
    print("%-58s %d iterations, %d launches, finite=%s" % (label, its, solver.last_fit_info.get("launches", -1), ok), flush=True)
