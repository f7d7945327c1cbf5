"""Parity at sizes the float64 CPU oracle cannot reach in seconds, through properties that do not depend on the size
(task section 3): the tensor-core engine against the engine's own float64 mode (itself pinned to the oracle at small
sizes), agreement between the three streamed-product code paths, exact scale covariance and permutation equivariance
of the iteration.  Relations are generated on the device by the counter-based generator (fz_fill_uniform)."""
import os

import numpy as np
import pytest

from helpers import rel_fro

pytestmark = pytest.mark.gpu

RANK = 64


def _graph(n, n_types, seed0=1000, dtype="bfloat16", sizes=None):
    import torch
    from skfusion import _capi
    tdt = {"bfloat16": torch.bfloat16, "float32": torch.float32, "float64": torch.float64}[dtype]
    types = list(range(n_types))
    sizes = sizes or {t: n for t in types}
    R = {}
    for i in types:
        for j in types:
            if i < j:
                t = torch.empty((sizes[i], sizes[j]), dtype=torch.bfloat16, device="cuda")
                _capi.fill_uniform(t, seed0 + 10 * i + j)
                R[i, j] = [t.to(tdt)]
    return types, sizes, {t: RANK for t in types}, R


def _fit(R, types, ranks, iters, seed=0, env=None, **opts):
    from skfusion.fusion import solver
    saved = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        return solver.dfmf(R, {}, types, ranks, max_iter=iters, init_type="random", random_state=np.random.RandomState(seed), **opts)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _worst(Ga, Sa, Gb, Sb):
    g = max(rel_fro(Ga[k], Gb[k]) for k in Ga)
    s = max(rel_fro(Sa[k][l], Sb[k][l]) for k in Sa for l in range(len(Sa[k])))
    return g, s


def test_tensor_core_engine_against_float64_engine_at_8192():
    """bf16 storage / 2-term tcgen05 path vs the float64 engine on the same (bf16-representable) numbers."""
    types, sizes, ranks, R = _graph(8192, 3)
    G, S = _fit(R, types, ranks, 8, dtype="float32", storage="bfloat16", split_terms=2)
    R64 = {k: [m.double() for m in v] for k, v in R.items()}
    G64, S64 = _fit(R64, types, ranks, 8, dtype="float64")
    g, s = _worst(G64, S64, G, S)
    assert g < 1e-3 and s < 5e-3, (g, s)           # the stated tolerance of the bf16 path; measured ~1e-5


@pytest.mark.parametrize("env", [{"FZ_NO_FUSED": "1"}, {"FZ_FUSED_VER": "4"}], ids=["two_pass", "fused_v4"])
def test_streamed_product_code_paths_agree(env):
    """The fused v3 kernel, the fused v4 kernel and the two-pass kernels compute the same products (ragged sizes)."""
    types, sizes, ranks, R = _graph(0, 3, sizes={0: 5000, 1: 3333, 2: 4100})
    Ga, Sa = _fit(R, types, ranks, 6, dtype="float32", storage="bfloat16")
    Gb, Sb = _fit(R, types, ranks, 6, env=env, dtype="float32", storage="bfloat16")
    g, s = _worst(Ga, Sa, Gb, Sb)
    assert g < 1e-4 and s < 1e-3, (g, s)           # same arithmetic, different summation orders (fp32); path tolerance is 1e-3 / 5e-3


def test_scale_covariance_on_the_100k_node_graph():
    """R -> 4 R leaves every G unchanged and multiplies every S by 4, exactly in exact arithmetic (S is linear in R, the
    update ratio is homogeneous of degree 0); a power of two keeps it exact in floating point up to the order of the
    L2 reductions.  5 types x 20 480 objects = the "100k-node graph" of BASELINE.json's metric (8.4 GB of bf16)."""
    types, sizes, ranks, R = _graph(20480, 5)
    G1, S1 = _fit(R, types, ranks, 4, dtype="float32", storage="bfloat16")
    for mats in R.values():
        mats[0].mul_(4.0)                           # exact in bf16
    G4, S4 = _fit(R, types, ranks, 4, dtype="float32", storage="bfloat16")
    S4 = {k: [m / 4.0 for m in v] for k, v in S4.items()}
    g, s = _worst(G1, S1, G4, S4)
    assert g < 5e-5 and s < 5e-4, (g, s)           # only the order of the L2 reductions differs between the two fits
    for k in G1:
        assert np.isfinite(G1[k]).all() and (G1[k] >= 0).all()


def test_permutation_equivariance():
    """Relabelling the objects of one type permutes the rows of its factor and nothing else: tiles, chunk boundaries and
    the wave-aware column split see different data, the result is the same up to summation order."""
    import torch
    types, sizes, ranks, R = _graph(0, 3, sizes={0: 4096, 1: 6000, 2: 2500})
    rs = np.random.RandomState(5)
    perm = rs.permutation(sizes[1])
    tperm = torch.from_numpy(perm).cuda()
    Rp = {}
    for (i, j), mats in R.items():
        m = mats[0]
        if i == 1:
            m = m.index_select(0, tperm)
        if j == 1:
            m = m.index_select(1, tperm)
        Rp[i, j] = [m.contiguous()]
    G0 = {t: np.random.RandomState(t).rand(sizes[t], RANK) for t in types}

    def run(Rx, G0x):
        # seed the factors explicitly: feed them through a callback-free fit with max_iter via the engine API
        from skfusion import _capi
        eng = _capi.Engine(0, "float32")
        tid = {t: eng.add_type(sizes[t], RANK) for t in types}
        rid = {key: eng.add_relation(tid[key[0]], tid[key[1]], mats[0], storage="bfloat16", borrow=mats[0].stride(0) % 8 == 0)
               for key, mats in Rx.items()}
        for t in types:
            eng.set_factor(tid[t], G0x[t])
        eng.finalize()
        eng.iterate(_capi.FZ_DFMF, 5)
        G = {t: eng.get_factor(tid[t]) for t in types}
        S = {key: eng.get_backbone(r) for key, r in rid.items()}
        eng.close()
        return G, S

    G, S = run(R, G0)
    G0p = dict(G0)
    G0p[1] = G0[1][perm]
    Gp, Sp = run(Rp, G0p)
    assert rel_fro(G[1][perm], Gp[1]) < 5e-5
    assert rel_fro(G[0], Gp[0]) < 5e-5 and rel_fro(G[2], Gp[2]) < 5e-5
    for key in S:
        assert rel_fro(S[key], Sp[key]) < 5e-4
