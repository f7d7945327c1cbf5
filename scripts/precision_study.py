"""Developer study (CPU, numpy): how the operand form of the factors in the two streamed products
A = R G_j and B = R^T G_i limits parity with the float64 oracle.  Emulates the engine's iteration (fp32 factors,
fp64 k x k chain, regrouped algebra) with configurable rounding of the factor operand per product.
    python scripts/precision_study.py [n] [iters]
"""
import os
import sys
import numpy as np
import scipy.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import fusion_oracle as oracle  # noqa: E402

EPS64 = np.finfo(float).eps


def bf16(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).astype(np.float64)


def fp16_scaled(x):
    """fp16 with an exact power-of-two scale per latent column (max -> [2^14, 2^15))."""
    x = np.asarray(x, dtype=np.float64)
    mx = np.maximum(np.abs(x).max(axis=0), 1e-300)
    e = 14 - np.floor(np.log2(mx))
    s = 2.0 ** e
    return (x * s).astype(np.float16).astype(np.float64) / s


def terms(x, n, rnd):
    out = np.zeros_like(np.asarray(x, dtype=np.float64))
    res = np.asarray(x, dtype=np.float64).copy()
    for _ in range(n):
        t = rnd(res)
        out += t
        res = res - t
    return out


SCHEMES = {
    "A bf16x2 / B bf16x2 (engine today)": (lambda g: terms(g, 2, bf16), lambda g: terms(g, 2, bf16)),
    "A bf16x2 / B bf16x1": (lambda g: terms(g, 2, bf16), lambda g: terms(g, 1, bf16)),
    "A bf16x2 / B fp16x1": (lambda g: terms(g, 2, bf16), lambda g: terms(g, 1, fp16_scaled)),
    "A fp16x2 / B fp16x1": (lambda g: terms(g, 2, fp16_scaled), lambda g: terms(g, 1, fp16_scaled)),
    "A fp16x1 / B fp16x1": (lambda g: terms(g, 1, fp16_scaled), lambda g: terms(g, 1, fp16_scaled)),
    "A bf16x1 / B bf16x1": (lambda g: terms(g, 1, bf16), lambda g: terms(g, 1, bf16)),
    "A fp32 / B fp32": (lambda g: g, lambda g: g),
}


def split(x):
    t = x > 0
    return t * x, (t - 1) * x


def emulate(R, types, ranks, G0, iters, fa, fb, Theta=None):
    G = {t: G0[t, t].astype(np.float32) for t in types}
    S = {}
    Theta = Theta or {}
    for _ in range(iters):
        g64 = {t: G[t].astype(np.float64) for t in types}
        gram = {t: np.nan_to_num(g64[t].T @ g64[t]) for t in types}
        P = {t: spla.pinv(gram[t]) for t in types}
        ga = {t: fa(g64[t]) for t in types}
        gb = {t: fb(g64[t]) for t in types}
        num = {t: np.zeros_like(g64[t]) for t in types}
        den = {t: np.zeros_like(g64[t]) for t in types}
        for (ti, tj), mats in R.items():
            for l, mat in enumerate(mats):
                A = (mat @ ga[tj]).astype(np.float32).astype(np.float64)
                B = (mat.T @ gb[ti]).astype(np.float32).astype(np.float64)
                M = np.nan_to_num(g64[ti].T @ A)
                Sij = np.nan_to_num(P[ti] @ M @ P[tj])
                S.setdefault((ti, tj), {})[l] = Sij
                t1p, t1n = split(np.nan_to_num((A @ Sij.T).astype(np.float32).astype(np.float64)))
                t2p, t2n = split(np.nan_to_num(Sij @ gram[tj] @ Sij.T))
                t4p, t4n = split(np.nan_to_num((B @ Sij).astype(np.float32).astype(np.float64)))
                t5p, t5n = split(np.nan_to_num(Sij.T @ gram[ti] @ Sij))
                num[ti] += t1p + g64[ti] @ t2n
                den[ti] += t1n + g64[ti] @ t2p
                num[tj] += t4p + g64[tj] @ t5n
                den[tj] += t4n + g64[tj] @ t5p
        for (t, _), mats in Theta.items():
            for th in mats:
                den[t] += np.maximum(th, 0) @ g64[t]
                num[t] += np.maximum(-th, 0) @ g64[t]
        for t in types:
            G[t] = (g64[t] * np.sqrt(num[t] / np.maximum(den[t], EPS64))).astype(np.float32)
    return {(t, t): G[t].astype(np.float64) for t in types}, {k: [d[l] for l in sorted(d)] for k, d in S.items()}


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(a)


def study(name, R, Theta, types, ranks, init, iters):
    import warnings
    warnings.simplefilter("ignore")
    sizes = oracle.count_objects(R)
    G0 = oracle.initialize(types, sizes, ranks, {k: v[0] for k, v in R.items()}, init, np.random.RandomState(0))
    Go, So = oracle.dfmf(R, Theta, types, ranks, max_iter=iters, G0=G0)
    conds = [np.linalg.cond(Go[t, t].T @ Go[t, t]) for t in types]
    print("== %s: init=%s iters=%d, cond(Gram) at the end: %s" % (name, init, iters, ", ".join("%.2g" % c for c in conds)))
    for label, (fa, fb) in SCHEMES.items():
        G, S = emulate(R, types, ranks, G0, iters, fa, fb, Theta)
        eg = max(rel(Go[t, t], G[t, t]) for t in types)
        es = max(rel(So[k][l], S[k][l]) for k in So for l in range(len(So[k])))
        print("   %-36s relFro(G) %.2e   relFro(S) %.2e" % (label, eg, es))


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 640
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    types, ranks, R = oracle.synthetic_graph(n, n_types=3, rank=64, storage="bfloat16")
    study("synthetic 3 types n=%d" % n, R, {}, types, ranks, "random", iters)
    study("synthetic 3 types n=%d" % n, R, {}, types, ranks, "random_c", iters)
    import cases
    c = cases.dicty_case()
    Rb = {k: [bf16(m) for m in v] for k, v in c["R"].items()}
    study("dicty (bf16-rounded relations)", Rb, c["Theta"], c["types"], c["ranks"], c["init_type"], iters)
