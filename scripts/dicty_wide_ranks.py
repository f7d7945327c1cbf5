"""dicty at the ranks of examples/dicty_factorization.py:37-40 (floor(0.7 n) = 853 / 81 / 197): the float64 engine and the
exact-planes tensor-core path (storage='bfloat16x3', constraint included) against the float64 oracle.  One JSON line each."""
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("scikit-fusion_b200", "oracle", os.path.join("tests", "golden")):
    sys.path.insert(0, os.path.join(ROOT, p))
import cases            # noqa: E402
import fusion_oracle as oracle   # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(a))


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    init = sys.argv[2] if len(sys.argv) > 2 else "random"
    c = cases.dicty_case()
    sizes = oracle.count_objects(c["R"])
    ranks = {t: int(0.7 * sizes[t]) for t in c["types"]}
    R = {k: [m.astype(np.float32).astype(np.float64) for m in v] for k, v in c["R"].items()}
    Th = {k: [m.astype(np.float32).astype(np.float64) for m in v] for k, v in c["Theta"].items()}
    warnings.simplefilter("ignore")
    t0 = time.time()
    Go, So = oracle.dfmf(R, Th, c["types"], ranks, max_iter=iters, init_type=init, random_state=np.random.RandomState(0))
    t_oracle = time.time() - t0
    print(json.dumps({"oracle_s": round(t_oracle, 2), "ranks": ranks, "iters": iters, "init": init,
                      "cond": [float(np.linalg.cond(Go[t, t].T @ Go[t, t])) for t in c["types"]]}), flush=True)
    if "--cpu" in sys.argv:
        return
    from skfusion.fusion import solver
    for dtype, storage in (("float64", None), ("float32", "bfloat16x3"), ("float32", None)):
        t0 = time.time()
        G, S = solver.dfmf(R, Th, c["types"], ranks, max_iter=iters, init_type=init, random_state=np.random.RandomState(0),
                           dtype=dtype, storage=storage)
        print(json.dumps({"dtype": dtype, "storage": storage, "fit_s": round(time.time() - t0, 2),
                          "relFro_G": max(rel(Go[t, t], G[t, t]) for t in c["types"]),
                          "relFro_S": max(rel(So[k][0], S[k][0]) for k in So), "launches": solver.last_fit_info.get("launches")}), flush=True)


if __name__ == "__main__":
    main()
