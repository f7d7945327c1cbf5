"""fz_objective (reference: the per-relation Frobenius residuals of _dfmf.py:306-319): the trace form
||R||^2 - 2 tr(S^T G_i^T R G_j) + tr(S^T G_i^T G_i S G_j^T G_j), which needs no n_i x n_j pass and hands its products to the next
iteration, against the oracle's explicit reconstruction; and the exact form it falls back to when the fit is nearly perfect."""
import numpy as np
import pytest

import fusion_oracle as oracle

pytestmark = pytest.mark.gpu


def _engine(R, types, ranks, G0, dtype, storage=None, terms=None):
    from skfusion import _capi
    eng = _capi.Engine(0, dtype)
    if terms is not None:
        eng.set_split_terms(terms)
    tid = {t: eng.add_type(G0[t, t].shape[0], ranks[t]) for t in types}
    rid = {key: eng.add_relation(tid[key[0]], tid[key[1]], mats[0], storage=storage) for key, mats in R.items()}
    for t in types:
        eng.set_factor(tid[t], G0[t, t])
    eng.finalize()
    return eng, tid, rid


@pytest.mark.parametrize("dtype,storage,terms,rtol", [("float64", None, None, 1e-9), ("float32", None, None, 2e-5),
                                                       ("float32", "bfloat16", "centred1", 1e-4), ("float32", "bfloat16", 2, 1e-4)])
def test_trace_form_objective_follows_the_oracle_and_feeds_the_next_iteration(dtype, storage, terms, rtol):
    from skfusion import _capi
    n = 700
    types, ranks, R = oracle.synthetic_graph(n, n_types=3, rank=32, storage=storage or "float64")
    sizes = {t: n for t in types}
    G0 = oracle.initialize(types, sizes, ranks, {}, "random", np.random.RandomState(0))
    hist = []
    oracle.dfmf(R, {}, types, ranks, max_iter=5, G0=G0, compute_err=True, history=hist)
    eng, tid, rid = _engine(R, types, ranks, G0, dtype, storage, terms)
    try:
        got = []
        for _ in range(5):
            eng.iterate(_capi.FZ_DFMF, 1)
            total, per = eng.objective(len(rid))
            assert abs(sum(per) - total) < 1e-9 * total
            got.append(total)
        np.testing.assert_allclose(got, hist, rtol=rtol)
        # interleaved objective calls must not change the trajectory ...
        G_a = eng.get_factor(tid[types[0]])
        eng2, tid2, _ = _engine(R, types, ranks, G0, dtype, storage, terms)
        eng2.iterate(_capi.FZ_DFMF, 5)
        G_b = eng2.get_factor(tid2[types[0]])
        assert np.abs(G_a - G_b).max() <= (1e-12 if dtype == "float64" else 2e-5) * np.abs(G_b).max()
        # ... and the products they ran are the next iteration's: objective + iterate costs no second pass over the relations
        base = eng2.launches
        eng2.iterate(_capi.FZ_DFMF, 1)
        plain = eng2.launches - base
        base = eng.launches
        eng.iterate(_capi.FZ_DFMF, 1)            # its products were run by the last objective call
        assert eng.launches - base < plain
        eng2.close()
    finally:
        eng.close()


def test_nearly_exact_fit_takes_the_explicit_form():
    """Full-rank factorisation of a small matrix: the residual is ~1e-7 of ||R||, far below what a difference of large numbers
    can resolve -- the objective then comes from the explicit n_i x n_j form and still matches the oracle."""
    from skfusion import _capi
    rs = np.random.RandomState(3)
    R = {("a", "b"): [rs.rand(40, 30)]}
    types, ranks = ["a", "b"], {"a": 40, "b": 30}
    G0 = oracle.initialize(types, {"a": 40, "b": 30}, ranks, {}, "random", np.random.RandomState(1))
    hist = []
    oracle.dfmf(R, {}, types, ranks, max_iter=3, G0=G0, compute_err=True, history=hist)
    eng, tid, rid = _engine(R, types, ranks, G0, "float64")
    try:
        eng.iterate(_capi.FZ_DFMF, 3)
        total, _ = eng.objective(1)
        norm_r = np.linalg.norm(R["a", "b"][0])
        assert hist[-1] < 1e-3 * norm_r
        assert abs(total - hist[-1]) <= 1e-9 * norm_r          # both are rounding noise around an exact fit
    finally:
        eng.close()
