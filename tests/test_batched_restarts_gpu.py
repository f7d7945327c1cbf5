"""Restarts batched two per pass over the relations (solver.dfmf_runs / fz_pair_iterate) against the same
restarts run one after the other and against the float64 oracle.  Reference semantics being kept: Dfmf(n_run=k) draws the k
initialisations from ONE RandomState in run order (decomposition/dfmf.py:63-64, 87-95 with n_jobs=1)."""
import warnings

import numpy as np
import pytest

import fusion_oracle as oracle
from helpers import rel_fro

pytestmark = pytest.mark.gpu


def _graph(n=1100):
    return oracle.synthetic_graph(n, n_types=3, rank=64, storage="bfloat16")


@pytest.mark.parametrize("init,terms", [("random", "centred1"), ("random_c", "centred1"), ("random", "auto")])
def test_batched_restarts_equal_sequential_restarts_and_follow_the_oracle(init, terms):
    from skfusion.fusion import solver
    types, ranks, R = _graph()
    iters, n_run = 8, 3
    kw = dict(dtype="float32", storage="bfloat16", split_terms=terms, device_init=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        runs = solver.dfmf_runs(R, {}, types, ranks, n_run, max_iter=iters, init_type=init, random_state=np.random.RandomState(7), **kw)
        info = dict(solver.last_fit_info)
        rs_seq, rs_o = np.random.RandomState(7), np.random.RandomState(7)
        seq = [solver.dfmf(R, {}, types, ranks, max_iter=iters, init_type=init, random_state=rs_seq, **kw) for _ in range(n_run)]
        ora = [oracle.dfmf(R, {}, types, ranks, max_iter=iters, init_type=init, random_state=rs_o) for _ in range(n_run)]
    assert len(runs) == n_run and info["batched_runs"] == n_run
    if terms == "centred1":
        assert info["operand_stats"]["paired"] == iters          # runs 0 and 1 went through the pair kernel every iteration
    for (G, S), (Gs, Ss), (Go, So) in zip(runs, seq, ora):
        for t in types:
            assert rel_fro(Gs[t, t], G[t, t]) < 2e-5             # same arithmetic up to summation order inside the MMAs / reductions
            assert rel_fro(Go[t, t], G[t, t]) < 1e-3
        for key in So:
            assert rel_fro(So[key][0], S[key][0]) < 5e-3


def test_estimator_batches_its_restarts_and_keeps_run_order():
    from skfusion import fusion
    rs = np.random.RandomState(0)
    a, b = fusion.ObjectType("a", 16), fusion.ObjectType("b", 24)
    rel = fusion.Relation(oracle.bf16_round(rs.rand(700, 900)), a, b)
    graph = fusion.FusionGraph([rel])
    kw = dict(max_iter=6, init_type="random", n_run=4, dtype="float32", storage="bfloat16", split_terms="centred1")
    batched = fusion.Dfmf(random_state=11, **kw).fuse(graph)
    plain = fusion.Dfmf(random_state=11, batch_runs=False, **kw).fuse(graph)
    fa, fp = list(batched.factor(a)), list(plain.factor(a))
    assert len(fa) == 4 and len(list(batched.backbone(rel))) == 4
    for run in range(4):
        assert rel_fro(fp[run], fa[run]) < 2e-5
        assert rel_fro(plain.backbone(rel, run), batched.backbone(rel, run)) < 1e-4
    assert rel_fro(fa[0], fa[1]) > 1e-2                          # different restarts, not copies
