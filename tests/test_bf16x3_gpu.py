"""GPU parity of the exact tensor-core form of float32 relations (storage='bfloat16x3', include/fz_fusion.h: FZ_BF16X3):
the relation is kept as up to three bf16 planes whose sum is the float32 value, and every plane goes through the same
tcgen05 kernels as a bf16-stored relation.  Covers what the bf16 storage cannot: data that bf16 would round, masked
relations (dfmc, _dfmc.py:319-325), constraint matrices (_dfmf.py:284-292), ranks above 64
(examples/dicty_factorization.py:37-40 uses ranks in the hundreds) and transform with constraints.

Tolerances (relative Frobenius error against the float64 oracle fed the SAME float32 values):
  G <= 1e-4, S <= 1e-3 on 'random' seeds -- an order of magnitude inside what bf16 STORAGE of this data can reach
  (6.6e-4 / 3.4e-3: the rounding of R itself), because here only the factor operand is rounded (two bf16 terms);
  G <= 1e-3, S <= 5e-3 (the stated tensor-core bar) on the ill-conditioned data-driven seeds.
"""
import warnings

import numpy as np
import pytest

import fusion_oracle as oracle
from helpers import rel_fro

pytestmark = pytest.mark.gpu


def _graph(ns, ranks, seed, integer=False):
    rs = np.random.RandomState(seed)
    types = ["t%d" % i for i in range(len(ns))]
    R = {}
    for i in range(len(ns)):
        for j in range(i + 1, len(ns)):
            m = rs.randint(0, 6, size=(ns[i], ns[j])).astype(np.float64) if integer else rs.rand(ns[i], ns[j])
            R[types[i], types[j]] = [m.astype(np.float32).astype(np.float64)]     # exactly float32-representable
    return types, dict(zip(types, ranks)), R


def _check(Go, So, G, S, types, tol_g, tol_s):
    for t in types:
        err = rel_fro(Go[t, t], G[t, t])
        assert err < tol_g, "G[%s] relFro=%.3g" % (t, err)
    for key in So:
        for l in range(len(So[key])):
            err = rel_fro(So[key][l], S[key][l])
            assert err < tol_s, "S%s[%d] relFro=%.3g" % (key, l, err)


@pytest.mark.parametrize("terms", [2, "auto"])
def test_float32_relations_on_the_tensor_cores_keep_every_bit_of_R(terms):
    """Fused kernels, one pass per plane (three planes: random float32 data).  'auto' = the centred operand form."""
    from skfusion.fusion import solver
    types, ranks, R = _graph((520, 392, 640), (40, 64, 24), 3)
    Go, So = oracle.dfmf(R, {}, types, ranks, max_iter=15, init_type="random", random_state=np.random.RandomState(0))
    G, S = solver.dfmf(R, {}, types, ranks, max_iter=15, init_type="random", random_state=np.random.RandomState(0),
                       dtype="float32", storage="bfloat16x3", split_terms=terms)
    _check(Go, So, G, S, types, 1e-4, 1e-3)
    info = dict(solver.last_fit_info)
    assert info["operand_stats"]["single"] + info["operand_stats"]["two_term"] == 15      # the tensor-core kernels ran


def test_exact_planes_beat_bf16_storage_on_data_bf16_would_round():
    from skfusion.fusion import solver
    types, ranks, R = _graph((520, 392, 640), (40, 64, 24), 3)
    Go, So = oracle.dfmf(R, {}, types, ranks, max_iter=15, init_type="random", random_state=np.random.RandomState(0))
    errs = {}
    for storage in ("bfloat16x3", "bfloat16"):
        G, S = solver.dfmf(R, {}, types, ranks, max_iter=15, init_type="random", random_state=np.random.RandomState(0),
                           dtype="float32", storage=storage, split_terms=2)
        errs[storage] = max(rel_fro(So[k][0], S[k][0]) for k in So)
    assert errs["bfloat16x3"] * 10 < errs["bfloat16"], errs


def test_integer_valued_relations_need_one_plane_only():
    """0..5 'ratings' are exact in one bf16 plane: same launches per iteration as bf16 storage, same result."""
    from skfusion.fusion import solver
    types, ranks, R = _graph((520, 392), (32, 48), 5, integer=True)
    out = {}
    for storage in ("bfloat16x3", "bfloat16"):
        G, S = solver.dfmf(R, {}, types, ranks, max_iter=6, init_type="random", random_state=np.random.RandomState(1),
                           dtype="float32", storage=storage, split_terms=2)
        out[storage] = (G, S, solver.last_fit_info["launches"])
    Go, So = oracle.dfmf(R, {}, types, ranks, max_iter=6, init_type="random", random_state=np.random.RandomState(1))
    _check(Go, So, out["bfloat16x3"][0], out["bfloat16x3"][1], types, 1e-4, 1e-3)
    # set-up differs by the two split kernels of the single relation; the six iterations launch the same kernels
    assert abs(out["bfloat16x3"][2] - out["bfloat16"][2]) <= 4, (out["bfloat16x3"][2], out["bfloat16"][2])


@pytest.mark.parametrize("dtype,storage,init", [("float32", "bfloat16x3", "random"), ("float32", "bfloat16x3", "random_vcol"),
                                                ("float32", "bfloat16", "random"), ("float64", None, "random")])
def test_ranks_above_64(dtype, storage, init):
    """Ranks 96 / 130 / 64: two-pass tensor-core kernels over 128-column blocks of the factor operand, and the k x k chain
    tiled through shared memory (csrc/fz_chain.cuh: block_mm); the float64 engine checks that chain to 1e-8."""
    from skfusion.fusion import solver
    types, ranks, R = _graph((520, 392, 300), (96, 130, 64), 3)
    if storage == "bfloat16":
        R = {k: [oracle.bf16_round(m) for m in v] for k, v in R.items()}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Go, So = oracle.dfmf(R, {}, types, ranks, max_iter=12, init_type=init, random_state=np.random.RandomState(0))
        G, S = solver.dfmf(R, {}, types, ranks, max_iter=12, init_type=init, random_state=np.random.RandomState(0),
                           dtype=dtype, storage=storage)
    tol = (1e-8, 1e-8) if dtype == "float64" else ((1e-4, 1e-3) if init == "random" else (1e-3, 5e-3))
    _check(Go, So, G, S, types, *tol)
    if dtype == "float64":      # the trace-form objective (fz_objective) runs the same tiled k x k products
        from skfusion import _capi
        from test_objective_gpu import _engine
        sizes = oracle.count_objects(R)
        G0 = oracle.initialize(types, sizes, ranks, {}, "random", np.random.RandomState(0))
        hist = []
        oracle.dfmf(R, {}, types, ranks, max_iter=3, G0=G0, compute_err=True, history=hist)
        eng, tid, rid = _engine(R, types, ranks, G0, dtype)
        try:
            got = []
            for _ in range(3):
                eng.iterate(_capi.FZ_DFMF, 1)
                got.append(eng.objective(len(rid))[0])
            np.testing.assert_allclose(got, hist, rtol=1e-9)
        finally:
            eng.close()


@pytest.mark.parametrize("terms", [2, "auto"])
def test_constraint_matrices_on_the_tensor_cores(terms):
    """Theta+ / Theta- as separate plane sets; 'auto' exercises the rank-1 part of the centred operand form."""
    from skfusion.fusion import solver
    types, ranks, R = _graph((520, 392, 300), (40, 64, 24), 7)
    import cases
    rs = np.random.RandomState(11)
    th0 = cases._sparse_sym_constraint(rs, 520, density=0.02, scale=0.1).astype(np.float32).astype(np.float64)    # both signs
    th1 = -np.where(rs.rand(392, 392) < 0.03, 0.005, 0.0).astype(np.float32).astype(np.float64)   # one sign only (like dicty's ppi): no Theta+ planes
    Theta = {("t0", "t0"): [th0], ("t1", "t1"): [th1, th1.T.copy()]}
    Go, So = oracle.dfmf(R, Theta, types, ranks, max_iter=12, init_type="random", random_state=np.random.RandomState(0))
    G, S = solver.dfmf(R, Theta, types, ranks, max_iter=12, init_type="random", random_state=np.random.RandomState(0),
                       dtype="float32", storage="bfloat16x3", split_terms=terms)
    _check(Go, So, G, S, types, 1e-4, 1e-3)
    # the constraints matter: without them the fit is measurably different
    G0, _ = oracle.dfmf(R, {}, types, ranks, max_iter=12, init_type="random", random_state=np.random.RandomState(0))
    assert rel_fro(G0["t0", "t0"], Go["t0", "t0"]) > 1e-2


def test_completion_with_masked_relations_on_the_tensor_cores():
    """dfmc: the unknown entries are re-imputed (and re-split into the planes) every iteration."""
    from skfusion.fusion import solver
    types, ranks, R = _graph((392, 520, 136), (24, 32, 8), 9)
    rs = np.random.RandomState(2)
    mask = rs.rand(392, 520) < 0.3
    M = {("t0", "t1"): [mask], ("t0", "t2"): [None], ("t1", "t2"): [None]}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Go, So = oracle.dfmc(R, M, {}, types, ranks, max_iter=12, init_type="random", random_state=np.random.RandomState(0))
        G, S = solver.dfmc(R, M, {}, types, ranks, max_iter=12, init_type="random", random_state=np.random.RandomState(0),
                           dtype="float32", storage="bfloat16x3")
    _check(Go, So, G, S, types, 2e-4, 2e-3)


def test_transform_with_a_constraint_on_the_tensor_cores():
    from skfusion.fusion import solver
    import cases
    types, ranks, R = _graph((520, 392, 300), (40, 64, 24), 13)
    Gf, Sf = oracle.dfmf(R, {}, types, ranks, max_iter=8, init_type="random", random_state=np.random.RandomState(0))
    tobj = {t: cases.Tag(t) for t in types}
    G = {(tobj[t], tobj[t]): Gf[t, t] for t in types}
    S = {(tobj[a], tobj[b]): [Sf[a, b][0]] for (a, b) in R}
    rs = np.random.RandomState(21)
    n_new = 264
    R_new = {(tobj["t0"], tobj["t1"]): [rs.rand(n_new, 392).astype(np.float32).astype(np.float64)],
             (tobj["t0"], tobj["t2"]): [rs.rand(n_new, 300).astype(np.float32).astype(np.float64)]}
    th = cases._sparse_sym_constraint(rs, n_new, density=0.03, scale=0.1).astype(np.float32).astype(np.float64)
    Th = {(tobj["t0"], tobj["t0"]): [th]}
    rk = {tobj[t]: r for t, r in ranks.items()}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = oracle.transform(R_new, Th, tobj["t0"], rk, G, S, max_iter=10, init_type="random",
                                random_state=np.random.RandomState(5))
        got = solver.transform(R_new, Th, tobj["t0"], rk, G, S, max_iter=10, init_type="random",
                               random_state=np.random.RandomState(5), dtype="float32", storage="bfloat16x3")
    assert rel_fro(want, got) < 1e-4


@pytest.mark.parametrize("terms", [2, "auto"])
def test_sharded_handles_with_exact_planes_match_the_oracle(terms):
    """Two shard handles on one GPU (host-spelled exchange, tests/test_sharded_engine_gpu.py): the plane passes accumulate
    into the same A and B partials."""
    from test_sharded_engine_gpu import _run
    graph = _graph((520, 392, 300), (40, 64, 24), 3)
    types, G, S, Go, So = _run(2, "bfloat16x3", "float32", None, 8, terms=terms, graph=graph)
    _check(Go, So, G, S, types, 1e-4, 1e-3)


@pytest.mark.parametrize("init", ["random_c", "random_vcol"])
def test_device_side_initialisation_streams_the_planes(init):
    """random_c / random_vcol with the column means computed on the GPU (device_init=True): the sampled means and the
    column norms come from the planes / the float32 master and equal the host initialisation."""
    from skfusion.fusion import solver
    types, ranks, R = _graph((520, 392, 300), (40, 64, 24), 17)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Go, So = oracle.dfmf(R, {}, types, ranks, max_iter=8, init_type=init, random_state=np.random.RandomState(3))
        G, S = solver.dfmf(R, {}, types, ranks, max_iter=8, init_type=init, random_state=np.random.RandomState(3),
                           dtype="float32", storage="bfloat16x3", device_init=True)
    _check(Go, So, G, S, types, 1e-3, 5e-3)


def test_estimators_take_the_storage_keyword():
    """Dfmf / Dfmc / DfmfTransform(storage='bfloat16x3') through the reference-facing API, constraint included."""
    import cases
    from skfusion.fusion import Dfmc, Dfmf, DfmfTransform, FusionGraph, ObjectType, Relation
    rs = np.random.RandomState(0)
    t1, t2, t3 = ObjectType("a", 24), ObjectType("b", 32), ObjectType("c", 16)
    R12 = rs.rand(392, 520).astype(np.float32).astype(np.float64)
    R13 = rs.rand(392, 264).astype(np.float32).astype(np.float64)
    th = cases._sparse_sym_constraint(rs, 392, density=0.02, scale=0.1)
    rels = [Relation(R12, t1, t2), Relation(R13, t1, t3), Relation(th, t1, t1)]
    kw = dict(max_iter=10, init_type="random")
    exact = Dfmf(random_state=np.random.RandomState(1), dtype="float64", **kw).fuse(FusionGraph(rels))
    fast = Dfmf(random_state=np.random.RandomState(1), dtype="float32", storage="bfloat16x3", **kw).fuse(FusionGraph(rels))
    for t in (t1, t2, t3):
        assert rel_fro(exact.factor(t), fast.factor(t)) < 1e-4
    assert rel_fro(exact.complete(rels[0]), fast.complete(rels[0])) < 1e-4
    # completion: 30 % of R12 unknown
    masked = np.ma.masked_array(R12, mask=rs.rand(392, 520) < 0.3)
    rels_c = [Relation(masked, t1, t2), Relation(R13, t1, t3)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        exact_c = Dfmc(random_state=np.random.RandomState(2), dtype="float64", **kw).fuse(FusionGraph(rels_c))
        fast_c = Dfmc(random_state=np.random.RandomState(2), dtype="float32", storage="bfloat16x3", **kw).fuse(FusionGraph(rels_c))
    assert rel_fro(exact_c.complete(rels_c[0]), fast_c.complete(rels_c[0])) < 2e-4
    # projection of new rows of type a
    new_graph = FusionGraph([Relation(R12[:136].copy(), t1, t2), Relation(R13[:136].copy(), t1, t3)])
    te = DfmfTransform(random_state=np.random.RandomState(3), dtype="float64", max_iter=10).transform(t1, new_graph, exact)
    tf = DfmfTransform(random_state=np.random.RandomState(3), dtype="float32", storage="bfloat16x3", max_iter=10).transform(t1, new_graph, exact)
    assert rel_fro(te.factor(t1), tf.factor(t1)) < 1e-4


def test_non_finite_entries_are_refused():
    from skfusion import _capi
    eng = _capi.Engine(device=0, compute="float32")
    try:
        a, b = eng.add_type(64, 8), eng.add_type(72, 8)
        bad = np.ones((64, 72), dtype=np.float32)
        bad[3, 4] = np.nan
        eng.add_relation(a, b, bad, storage="bfloat16x3")
        with pytest.raises(_capi.EngineError):
            eng.finalize()
    finally:
        eng.close()
