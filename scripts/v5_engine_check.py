"""GPU-box check: the centred operand forms through the engine against the float64 oracle (single- and two-term kernels)."""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-fusion_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fusion_oracle as oracle
from skfusion.fusion import solver
warnings.simplefilter("ignore")
def rel(a, b): return np.linalg.norm(a - b) / np.linalg.norm(a)
for n, nt, init, iters in ((384, 3, "random", 8), (1000, 3, "random", 12), (1280, 3, "random_c", 12), (2048, 4, "random", 10), (1280, 3, "random_vcol", 12)):
    types, ranks, R = oracle.synthetic_graph(n, n_types=nt, rank=64, storage="bfloat16")
    Go, So = oracle.dfmf(R, {}, types, ranks, max_iter=iters, init_type=init, random_state=np.random.RandomState(0))
    for mode in (2, "auto", "centred1"):
        G, S = solver.dfmf(R, {}, types, ranks, max_iter=iters, init_type=init, random_state=np.random.RandomState(0),
                           dtype="float32", storage="bfloat16", split_terms=mode, device_init=False)
        eg = max(rel(Go[t, t], G[t, t]) for t in types)
        es = max(rel(So[k][l], S[k][l]) for k in So for l in range(len(So[k])))
        print("n=%d types=%d init=%s iters=%d split_terms=%s : relFro(G) %.2e relFro(S) %.2e" % (n, nt, init, iters, mode, eg, es), flush=True)
