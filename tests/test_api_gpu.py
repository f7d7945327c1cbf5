"""The reference's own API-level scenarios (skfusion/tests/test_dfmf.py, test_dfmc.py, test_base.py,
test_multiple_relations.py, test_n_run.py) replayed on the CUDA engine through skfusion.fusion.
float64 engine: the reference's 7-decimal full-rank assertions hold as written (F9).
float32 engine: same properties at fp32-level tolerance (abs 2e-4), stated per test."""
import warnings

import numpy as np
import pytest

from skfusion.fusion import Dfmc, Dfmf, DfmfTransform, FusionGraph, ObjectType, Relation

pytestmark = pytest.mark.gpu

MODES = [("float64", 7), ("float32", 3)]


@pytest.mark.parametrize("dtype,decimal", MODES)
def test_dfmf_full_rank_reconstruction(dtype, decimal):           # test_dfmf.py:9-23
    rnds = np.random.RandomState(0)
    R12 = rnds.rand(50, 30)
    t1, t2 = ObjectType('type1', 50), ObjectType('type2', 30)
    relation = Relation(R12, t1, t2)
    fuser = Dfmf(init_type='random', random_state=rnds, dtype=dtype).fuse(FusionGraph([relation]))
    assert fuser.backbone(relation).shape == (50, 30)
    assert fuser.factor(t1).shape == (50, 50) and fuser.factor(t2).shape == (30, 30)
    np.testing.assert_almost_equal(fuser.complete(relation), relation.data, decimal=decimal)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_infinite_inputs_give_finite_completion(dtype):           # test_dfmf.py:25-46
    rnds = np.random.RandomState(0)
    R12 = rnds.rand(50, 30)
    R13 = rnds.rand(50, 10)
    R12 = np.ma.masked_greater(R12, 0.7)
    R12[R12 < 0.1] = np.nan
    R13[R13 < 0.5] = np.inf
    t1, t2, t3 = ObjectType('type1', 50), ObjectType('type2', 30), ObjectType('type3', 10)
    relations = [Relation(R12, t1, t2, fill_value='row_mean'), Relation(R13, t1, t3, fill_value='col_mean')]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fuser = Dfmf(init_type='random', random_state=rnds, dtype=dtype).fuse(FusionGraph(relations))
        assert fuser.backbone(relations[0]).shape == (50, 30) and fuser.backbone(relations[1]).shape == (50, 10)
        assert np.sum(np.isfinite(fuser.complete(relations[0]))) == R12.size


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_transformation_recovers_training_rows(dtype):            # test_dfmf.py:48-78
    R12 = np.random.RandomState(42).rand(5, 3)
    t1, t2 = ObjectType('type1', 2), ObjectType('type2', 2)
    relation = Relation(R12, t1, t2)
    fuser = Dfmf(init_type='random', random_state=np.random.RandomState(0), max_iter=100, dtype=dtype).fuse(
        FusionGraph([relation]))
    new_graph = FusionGraph([Relation(R12[:2].copy(), t1, t2)])
    transformer = DfmfTransform(random_state=np.random.RandomState(0), dtype=dtype).transform(t1, new_graph, fuser)
    new_G1, G1, G2, S12 = transformer.factor(t1), fuser.factor(t1), fuser.factor(t2), fuser.backbone(relation)
    diff_G1 = new_G1 - G1[:2]
    diff_hat = new_G1 @ S12 @ G2.T - (G1 @ S12 @ G2.T)[:2]
    assert np.sum(diff_G1 ** 2) / diff_G1.size < 1e-5
    assert np.sum(diff_hat ** 2) / diff_hat.size < 1e-5


@pytest.mark.parametrize("dtype,decimal", MODES)
def test_pre_and_postprocessors(dtype, decimal):                  # test_dfmf.py:80-120
    rnds = np.random.RandomState(0)
    R12 = rnds.rand(50, 30)
    t1, t2 = ObjectType('type1', 50), ObjectType('type2', 30)
    pre = Relation(R12, t1, t2, preprocessor=lambda d: np.ones_like(d))
    fuser = Dfmf(init_type='random', random_state=rnds, dtype=dtype).fuse(FusionGraph([pre]))
    np.testing.assert_almost_equal(fuser.complete(pre), np.ones_like(R12), decimal=decimal)
    post = Relation(R12, t1, t2, postprocessor=lambda d: d - np.mean(d))
    fuser = Dfmf(init_type='random', random_state=rnds, dtype=dtype).fuse(FusionGraph([post]))
    np.testing.assert_almost_equal(fuser.complete(post), R12 - np.mean(R12), decimal=decimal)


@pytest.mark.parametrize("dtype,decimal", MODES)
def test_dfmc_full_rank_and_masked(dtype, decimal):               # test_dfmc.py:9-39
    rnds = np.random.RandomState(0)
    R12 = rnds.rand(50, 30)
    t1, t2 = ObjectType('type1', 50), ObjectType('type2', 30)
    relation = Relation(R12, t1, t2)
    fuser = Dfmc(init_type='random', random_state=rnds, dtype=dtype).fuse(FusionGraph([relation]))
    np.testing.assert_almost_equal(fuser.complete(relation), R12, decimal=decimal)
    Rm = np.ma.masked_greater(rnds.rand(50, 30), 0.8)
    before = Rm.copy()
    relation = Relation(Rm, t1, t2)
    fuser = Dfmc(init_type='random', random_state=rnds, dtype=dtype).fuse(FusionGraph([relation]))
    completed = fuser.complete(relation)
    np.testing.assert_almost_equal(completed[~Rm.mask], Rm.data[~Rm.mask], decimal=decimal)
    np.testing.assert_array_equal(Rm.data, before.data)           # input not mutated (test_dfmc.py:62)
    np.testing.assert_array_equal(Rm.mask, before.mask)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_pipeline_shapes_with_rank_above_object_count(dtype):     # test_base.py:9-37 (default random_c init)
    rnds = np.random.RandomState(0)
    R12, R13, R23 = rnds.rand(50, 30), rnds.rand(50, 40), rnds.rand(30, 40)
    t1, t2, t3 = ObjectType('type1', 30), ObjectType('type2', 40), ObjectType('type3', 40)
    relations = [Relation(R12, t1, t2), Relation(R13, t1, t3), Relation(R23, t2, t3)]
    fuser = Dfmf(random_state=rnds, dtype=dtype).fuse(FusionGraph(relations))
    assert fuser.factor(t1).shape == (50, 30) and fuser.factor(t2).shape == (30, 40) and fuser.factor(t3).shape == (40, 40)
    assert fuser.backbone(relations[0]).shape == (30, 40) and fuser.backbone(relations[2]).shape == (40, 40)
    assert all(np.isfinite(fuser.factor(t)).all() for t in (t1, t2, t3))
    new_graph = FusionGraph([Relation(rnds.rand(15, 30), t1, t2), Relation(rnds.rand(15, 40), t1, t3)])
    transformer = DfmfTransform(random_state=rnds, dtype=dtype).transform(t1, new_graph, fuser)
    assert transformer.factor(t1).shape == (15, 30) and np.isfinite(transformer.factor(t1)).all()


@pytest.mark.parametrize("cls", [Dfmf, Dfmc])
def test_multiple_relations_and_runs(cls):                        # test_multiple_relations.py, test_n_run.py
    rnds = np.random.RandomState(0)
    t1, t2, t3 = ObjectType('type1', 30), ObjectType('type2', 30), ObjectType('type3', 20)
    relations = [Relation(rnds.rand(30, 30), t1, t2), Relation(rnds.rand(30, 30), t1, t2), Relation(rnds.rand(30, 20), t1, t3)]
    fuser = cls(init_type='random', random_state=rnds, n_run=3, max_iter=20).fuse(FusionGraph(relations))
    assert len(list(fuser.factor(t1))) == 3 and len(list(fuser.backbone(relations[0]))) == 3
    for run in range(3):
        G1, G2 = fuser.factor(t1, run), fuser.factor(t2, run)
        for rel in relations[:2]:
            np.testing.assert_almost_equal(fuser.complete(rel, run), G1 @ fuser.backbone(rel, run) @ G2.T)
    assert not np.allclose(fuser.backbone(relations[0], 0), fuser.backbone(relations[1], 0))
