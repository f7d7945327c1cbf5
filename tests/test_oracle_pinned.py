"""Pin the oracle against the REAL reference, live (build container only: needs /root/reference; skipped elsewhere,
where tests/test_oracle_golden.py checks the same thing against the committed fixtures).  The reference is imported
in memory with four py3.12 / numpy-2 compatibility edits (tests/golden/_load_reference.py); nothing is copied."""
import warnings

import numpy as np
import pytest

import _load_reference as ref
import cases
import fusion_oracle as oracle
from helpers import rel_fro

pytestmark = pytest.mark.skipif(not ref.available(), reason="reference sources not present on this box")


@pytest.mark.parametrize("init_type", ["random", "random_c", "random_vcol"])
@pytest.mark.parametrize("algo", ["dfmf", "dfmc"])
def test_oracle_equals_reference_free_functions(algo, init_type):
    r_dfmf, r_dfmc, _, _ = ref.functions()
    case = cases.fit_cases()["completion" if algo == "dfmc" else "multi_theta"]
    kw = dict(obj_types=case["types"], obj_type2rank=case["ranks"], max_iter=25, init_type=init_type)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if algo == "dfmc":
            G1, S1 = r_dfmc(case["R"], case["M"], case["Theta"], random_state=np.random.RandomState(4), **kw)
            G2, S2 = oracle.dfmc(case["R"], case["M"], case["Theta"], random_state=np.random.RandomState(4), **kw)
        else:
            G1, S1 = r_dfmf(case["R"], case["Theta"], random_state=np.random.RandomState(4), **kw)
            G2, S2 = oracle.dfmf(case["R"], case["Theta"], random_state=np.random.RandomState(4), **kw)
    # (the unconstrained 'random' start diverges on this graph -- in both, identically: compare element-wise)
    for key in G1:
        np.testing.assert_allclose(G2[key], G1[key], rtol=1e-11, atol=0, equal_nan=True)
    for key in S1:
        for a, b in zip(S1[key], S2[key]):
            np.testing.assert_allclose(b, a, rtol=1e-9, atol=1e-300, equal_nan=True)


def test_oracle_equals_reference_transform():
    _, _, r_transform, _ = ref.functions()
    fit = cases.fit_cases()["readme3"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G, S = oracle.dfmf(fit["R"], {}, fit["types"], fit["ranks"], max_iter=10, init_type="random",
                           random_state=np.random.RandomState(0))
    tags = {t: cases.Tag(t) for t in fit["types"]}
    Gd = {(tags[t], tags[t]): G[t, t] for t in fit["types"]}
    Sd = {(tags[a], tags[b]): v for (a, b), v in S.items()}
    tc = cases.transform_cases()["project_cols"]
    R_new = {(tags[a], tags[b]): m for (a, b), m in tc["R_new"].items()}
    Th = {(tags[a], tags[a]): m for (a, _), m in tc["Theta"].items()}
    rk = {tags[t]: r for t, r in fit["ranks"].items()}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g1 = r_transform(R_new, Th, tags[tc["target"]], rk, Gd, Sd, max_iter=30, init_type="random_c", random_state=np.random.RandomState(2))
        g2 = oracle.transform(R_new, Th, tags[tc["target"]], rk, Gd, Sd, max_iter=30, init_type="random_c", random_state=np.random.RandomState(2))
    assert rel_fro(g1, g2) < 1e-12


def test_host_layer_matches_reference_estimators_in_process():
    """Same process, same set-iteration order: the reference's Dfmf class and ours (with the oracle as
    solver backend) must consume the RNG identically and agree on every factor."""
    import skfusion.fusion as mine
    from skfusion.fusion import solver
    theirs = ref.load()
    rs = np.random.RandomState(0)
    R12, R13, R23 = rs.rand(50, 30), rs.rand(50, 40), rs.rand(30, 40)

    def graph(mod):
        t1, t2, t3 = mod.ObjectType('type1', 6), mod.ObjectType('type2', 8), mod.ObjectType('type3', 5)
        rels = [mod.Relation(R12, t1, t2), mod.Relation(R13, t1, t3), mod.Relation(R23, t2, t3)]
        return mod.FusionGraph(rels), (t1, t2, t3), rels
    g_ref, types_ref, rels_ref = graph(theirs)
    g_own, types_own, rels_own = graph(mine)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        f_ref = theirs.Dfmf(max_iter=15, random_state=7).fuse(g_ref)
        saved = solver.dfmf
        solver.dfmf = lambda **kw: oracle.dfmf(**{k: v for k, v in kw.items() if k not in ("device", "dtype", "storage", "split_terms")})
        try:
            f_own = mine.Dfmf(max_iter=15, random_state=7).fuse(g_own)
        finally:
            solver.dfmf = saved
    for a, b in zip(types_ref, types_own):
        assert rel_fro(f_ref.factor(a), f_own.factor(b)) < 1e-12
    for a, b in zip(rels_ref, rels_own):
        assert rel_fro(f_ref.backbone(a), f_own.backbone(b)) < 1e-11
        assert rel_fro(f_ref.complete(a), f_own.complete(b)) < 1e-11


@pytest.mark.parametrize("mode", ["mean", "row_mean", "col_mean", "const"])
@pytest.mark.parametrize("masked", [False, True])
def test_fill_functions_equal_the_reference(mode, masked):
    """Relation.filled(): our host fill functions against the reference's (fusion_graph.py:464-510) on NaN / masked data.
    The device fill (fz_fill_unknown) is checked against ours in tests/test_device_preprocessing.py, which closes the chain."""
    from skfusion.fusion import graph as mine
    theirs = ref.load()
    import importlib
    fg = importlib.import_module(theirs.FusionGraph.__module__)
    rs = np.random.RandomState(3)
    x = rs.rand(23, 17) * 5 - 1
    x[rs.rand(23, 17) < 0.2] = np.nan
    x[4, :] = np.nan
    if masked:
        x = np.ma.masked_array(np.nan_to_num(x, nan=7.0), mask=np.isnan(x) | (rs.rand(23, 17) < 0.1))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if mode == "const":
            want, got = fg.fill_const(x, 0.5), mine.fill_const(x, 0.5)
        else:
            want, got = fg.FILL_TYPE[mode](x), mine.FILL_TYPE[mode](x)
    assert np.ma.is_masked(want) == np.ma.is_masked(got)
    np.testing.assert_array_equal(np.ma.getdata(want), np.ma.getdata(got))
    if np.ma.is_masked(want):
        np.testing.assert_array_equal(np.ma.getmaskarray(want), np.ma.getmaskarray(got))


def _strip_engine_kwargs(fn):
    return lambda **kw: fn(**{k: v for k, v in kw.items() if k not in ("device", "dtype", "storage", "split_terms", "device_init")})


def test_host_layer_matches_reference_dfmc_and_transform_in_process():
    """Estimator level, same process: Dfmc with masked entries and two fill modes, then DfmfTransform of new
    rows against the fitted model -- marshalling (filled / preprocessor / mask survival, dfmc.py:69-94; dfmf.py:175-189),
    RNG order and accessors must match the reference's classes when the oracle stands in for the engine."""
    import skfusion.fusion as mine
    from skfusion.fusion import solver
    theirs = ref.load()
    rs = np.random.RandomState(1)
    R12 = np.ma.masked_array(rs.rand(40, 30), mask=rs.rand(40, 30) < 0.25)
    R13 = rs.rand(40, 25)
    R13[rs.rand(40, 25) < 0.1] = np.nan
    R23 = rs.rand(30, 25)
    T1 = (rs.rand(40, 40) < 0.1) * -0.05
    T1 = (T1 + T1.T) / 2
    R12_new, R13_new = rs.rand(12, 30), rs.rand(12, 25)

    def build(mod):
        t1, t2, t3 = mod.ObjectType('type1', 6), mod.ObjectType('type2', 5), mod.ObjectType('type3', 4)
        rels = [mod.Relation(R12.copy(), t1, t2), mod.Relation(R13.copy(), t1, t3, fill_value='row_mean'),
                mod.Relation(R23.copy(), t2, t3, preprocessor=lambda x: x * 2.0, postprocessor=lambda x: x / 2.0),
                mod.Relation(T1.copy(), t1, t1)]
        new = [mod.Relation(R12_new.copy(), t1, t2), mod.Relation(R13_new.copy(), t1, t3)]
        return mod.FusionGraph(rels), mod.FusionGraph(new), (t1, t2, t3), rels

    out = {}
    saved = (solver.dfmc, solver.transform)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name, mod in (("ref", theirs), ("own", mine)):
            g, g_new, types, rels = build(mod)
            if name == "own":
                solver.dfmc, solver.transform = _strip_engine_kwargs(oracle.dfmc), _strip_engine_kwargs(oracle.transform)
            try:
                fuser = mod.Dfmc(max_iter=12, init_type='random_vcol', random_state=3).fuse(g)
                tr = mod.DfmfTransform(max_iter=15, random_state=5).transform(types[0], g_new, fuser)
            finally:
                solver.dfmc, solver.transform = saved
            out[name] = ([fuser.factor(t) for t in types], [fuser.backbone(r) for r in rels[:3]],
                         [fuser.complete(r) for r in rels[:3]], tr.factor(types[0]))
    for a, b in zip(out["ref"][0], out["own"][0]):
        assert rel_fro(a, b) < 1e-11
    for a, b in zip(out["ref"][1], out["own"][1]):
        assert rel_fro(a, b) < 1e-10
    for a, b in zip(out["ref"][2], out["own"][2]):
        assert rel_fro(a, b) < 1e-10
    assert rel_fro(out["ref"][3], out["own"][3]) < 1e-11


@pytest.mark.parametrize("which", ["wide_ranks", "constraints", "completion"])
def test_oracle_equals_reference_on_the_graphs_of_the_exact_planes_tests(which):
    """tests/test_bf16x3_gpu.py compares the tensor-core paths with the oracle on graphs of a few hundred objects with ranks
    above 64, constraint matrices of both signs and a masked relation: pin the oracle to the real reference on those very
    inputs (same seeds, fewer iterations)."""
    import test_bf16x3_gpu as x3
    r_dfmf, r_dfmc, _, _ = ref.functions()
    Theta, M = {}, None
    if which == "wide_ranks":
        types, ranks, R = x3._graph((520, 392, 300), (96, 130, 64), 3)
    elif which == "constraints":
        types, ranks, R = x3._graph((520, 392, 300), (40, 64, 24), 7)
        rs = np.random.RandomState(11)
        th0 = cases._sparse_sym_constraint(rs, 520, density=0.02, scale=0.1).astype(np.float32).astype(np.float64)
        th1 = -np.where(rs.rand(392, 392) < 0.03, 0.005, 0.0).astype(np.float32).astype(np.float64)
        Theta = {("t0", "t0"): [th0], ("t1", "t1"): [th1, th1.T.copy()]}
    else:
        types, ranks, R = x3._graph((392, 520, 136), (24, 32, 8), 9)
        M = {("t0", "t1"): [np.random.RandomState(2).rand(392, 520) < 0.3], ("t0", "t2"): [None], ("t1", "t2"): [None]}
    kw = dict(obj_types=types, obj_type2rank=ranks, max_iter=4, init_type="random")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if M is not None:
            G1, S1 = r_dfmc(R, M, Theta, random_state=np.random.RandomState(0), **kw)
            G2, S2 = oracle.dfmc(R, M, Theta, random_state=np.random.RandomState(0), **kw)
        else:
            G1, S1 = r_dfmf(R, Theta, random_state=np.random.RandomState(0), **kw)
            G2, S2 = oracle.dfmf(R, Theta, random_state=np.random.RandomState(0), **kw)
    for key in G1:
        assert rel_fro(G1[key], G2[key]) < 1e-11
    for key in S1:
        for a, b in zip(S1[key], S2[key]):
            assert rel_fro(a, b) < 1e-9
