"""Host layer of the drop-in (no GPU): RNG-exact initialisation, graph -> block marshalling, accessors,
n_run bookkeeping.  The seam functions are swapped for the float64 oracle so the estimator classes can be
exercised on a CPU box; tests/test_api_gpu.py runs the same scenarios on the real engine."""
import warnings

import numpy as np
import pytest

import cases
import fusion_oracle as oracle
from skfusion.fusion import Dfmc, Dfmf, DfmfTransform, FusionGraph, ObjectType, Relation
from skfusion.fusion import initializers, solver
from skfusion.fusion.graph import DataFusionError


@pytest.fixture
def oracle_backend(monkeypatch):
    def strip(kw):
        return {k: v for k, v in kw.items() if k not in ("device", "dtype", "storage", "split_terms")}
    monkeypatch.setattr(solver, "dfmf", lambda **kw: oracle.dfmf(**strip(kw)))
    monkeypatch.setattr(solver, "dfmc", lambda **kw: oracle.dfmc(**strip(kw)))
    monkeypatch.setattr(solver, "transform", lambda **kw: oracle.transform(**strip(kw)))


@pytest.mark.parametrize("name", list(cases.fit_cases().keys()))
def test_initializers_consume_the_rng_like_the_reference(golden, name):
    case = cases.fit_cases()[name]
    sizes = solver.count_objects(case["types"], case["R"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G0 = initializers.initialize(case["types"], sizes, case["ranks"], {k: v[0] for k, v in case["R"].items()},
                                     case["init_type"], np.random.RandomState(case["seed"]))
    for t in case["types"]:
        np.testing.assert_array_equal(G0[t, t], golden["%s/G0/%s" % (name, t)])


def test_full_rank_fit_reconstructs_exactly(oracle_backend):
    rnds = np.random.RandomState(0)
    R12 = rnds.rand(50, 30)
    t1, t2 = ObjectType('type1', 50), ObjectType('type2', 30)
    rel = Relation(R12, t1, t2)
    fuser = Dfmf(init_type='random', random_state=rnds).fuse(FusionGraph([rel]))
    assert fuser.backbone(rel).shape == (50, 30) and fuser.factor(t1).shape == (50, 50)
    np.testing.assert_almost_equal(fuser.complete(rel), R12)


def test_marshalling_fill_preprocess_and_masks(oracle_backend, monkeypatch):
    rnds = np.random.RandomState(0)
    R12 = np.ma.masked_greater(rnds.rand(20, 15), 0.7)
    R13 = rnds.rand(20, 10)
    R13[R13 < 0.2] = np.nan
    th = rnds.rand(20, 20)
    t1, t2, t3 = ObjectType('a', 4), ObjectType('b', 3), ObjectType('c', 2)
    rels = [Relation(R12, t1, t2), Relation(R13, t1, t3, fill_value='row_mean', preprocessor=lambda d: d * 2),
            Relation(th, t1, t1)]
    seen = {}

    def spy(**kw):
        seen.update(kw)
        return oracle.dfmc(**{k: v for k, v in kw.items() if k not in ("device", "dtype", "storage", "split_terms")})
    monkeypatch.setattr(solver, "dfmc", spy)
    data_before = R12.copy()
    Dfmc(init_type='random', random_state=1, max_iter=3, dtype='float64').fuse(FusionGraph(rels))
    assert set(seen["R"].keys()) == {(t1, t2), (t1, t3)} and list(seen["Theta"].keys()) == [(t1, t1)]
    assert seen["M"][t1, t2][0] is not None and seen["M"][t1, t2][0].sum() == R12.mask.sum()
    assert seen["M"][t1, t3] == [None]
    assert not np.ma.isMaskedArray(seen["R"][t1, t2][0]) and np.isfinite(seen["R"][t1, t3][0]).all()
    np.testing.assert_allclose(seen["R"][t1, t3][0][~np.isnan(R13)], 2 * R13[~np.isnan(R13)])
    assert seen["dtype"] == 'float64' and isinstance(seen["obj_types"], set)
    np.testing.assert_array_equal(R12.data, data_before.data)      # inputs are never mutated
    np.testing.assert_array_equal(R12.mask, data_before.mask)


def test_multiple_relations_and_n_run(oracle_backend):
    rnds = np.random.RandomState(0)
    t1, t2, t3 = ObjectType('type1', 30), ObjectType('type2', 30), ObjectType('type3', 20)
    rels = [Relation(rnds.rand(30, 30), t1, t2), Relation(rnds.rand(30, 30), t1, t2), Relation(rnds.rand(30, 20), t1, t3)]
    fuser = Dfmf(init_type='random', random_state=rnds, n_run=3, max_iter=10).fuse(FusionGraph(rels))
    runs = list(fuser.factor(t1))
    assert len(runs) == 3 and not np.allclose(runs[0], runs[1])    # one RNG shared sequentially by the restarts
    assert fuser.factor(t1, run=2).shape == (30, 30)
    S0, S1 = fuser.backbone(rels[0], run=1), fuser.backbone(rels[1], run=1)
    assert S0.shape == (30, 30) and not np.allclose(S0, S1)
    G1, G2 = fuser.factor(t1, 1), fuser.factor(t2, 1)
    for rec, S in zip((fuser.complete(rels[0], run=1), fuser.complete(rels[1], run=1)), (S0, S1)):
        np.testing.assert_almost_equal(rec, G1 @ S @ G2.T)
    assert len(list(fuser.complete(rels[2]))) == 3 and len(list(fuser.backbone(rels[2]))) == 3


def test_transform_pipeline_and_validation(oracle_backend):
    R12 = np.random.RandomState(3).rand(5, 3)
    t1, t2 = ObjectType('type1', 2), ObjectType('type2', 2)
    rel = Relation(R12, t1, t2)
    fuser = Dfmf(init_type='random', random_state=np.random.RandomState(0), max_iter=100).fuse(FusionGraph([rel]))
    new_graph = FusionGraph([Relation(R12[:2].copy(), t1, t2)])
    tr = DfmfTransform(random_state=np.random.RandomState(0)).transform(t1, new_graph, fuser)
    new_G1 = tr.factor(t1)
    assert new_G1.shape == (2, 2)
    diff = new_G1 - fuser.factor(t1)[:2]
    assert np.sum(diff ** 2) / diff.size < 1e-5
    with pytest.raises(DataFusionError):
        DfmfTransform().transform(ObjectType('other'), new_graph, fuser)
    bad = FusionGraph([Relation(R12[:2].copy(), t1, t2), Relation(np.ones((3, 3)), t2, t2)])
    with pytest.raises(DataFusionError):
        DfmfTransform().transform(t1, bad, fuser)


def test_accessor_errors_and_repr(oracle_backend):
    t1, t2 = ObjectType('a', 2), ObjectType('b', 2)
    rel = Relation(np.random.RandomState(0).rand(6, 5), t1, t2)
    fuser = Dfmf(max_iter=2, init_type='random', random_state=0).fuse(FusionGraph([rel]))
    with pytest.raises(DataFusionError):
        fuser.factor(ObjectType('zzz'))
    with pytest.raises(DataFusionError):
        fuser.backbone(Relation(np.zeros((6, 5)), t1, t2))
    assert repr(fuser).startswith('Dfmf(max_iter=2, init_type=random')
    with pytest.raises(TypeError):
        Dfmf(bogus=1)
    assert list(fuser.chain(t1, t2)) == [[t1, t2]] and list(fuser.chain(t1, t1)) == [[t1]]
