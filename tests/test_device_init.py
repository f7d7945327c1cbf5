"""Factor initialisation with the column means on the GPU (reference _init.py:20-61).

CPU part: the sampling plan drawn on the host (initializers.sample_plan) reproduces the reference's selections and
leaves the RandomState in the reference's state -- checked against the host initializer (itself pinned bit-exactly to
the reference's by tests/test_oracle_pinned.py / test_host_layer.py).
GPU part: fz_init_* against the host initializer on the same seeds; a whole fit seeded on the device against the oracle.
"""
import warnings

import numpy as np
import pytest

import fusion_oracle as oracle
from helpers import rel_fro
from skfusion.fusion import initializers


def _graph(seed=0, sizes=(37, 52, 23), ranks=(4, 7, 5)):
    rs = np.random.RandomState(seed)
    types = ["a", "b", "c"]
    n = dict(zip(types, sizes))
    k = dict(zip(types, ranks))
    R = {("a", "b"): [rs.rand(n["a"], n["b"])], ("b", "c"): [rs.rand(n["b"], n["c"]) - 0.3],
         ("a", "c"): [rs.rand(n["a"], n["c"]), rs.rand(n["a"], n["c"])]}
    return types, n, k, R


@pytest.mark.parametrize("init_type", ["random_vcol", "random_c"])
def test_sample_plan_reproduces_the_host_initializer(init_type):
    types, n, k, R = _graph()
    first = {key: mats[0] for key, mats in R.items()}
    rs_ref, rs_plan = np.random.RandomState(11), np.random.RandomState(11)
    want = initializers.initialize(types, n, k, first, init_type, rs_ref)
    for t in types:
        total = 1e-5 * np.ones((n[t], k[t]))
        for pair, mat in first.items():
            if t not in pair:
                continue
            view = mat if t == pair[0] else mat.T
            norms = [np.linalg.norm(view[:, c], 2) for c in range(view.shape[1])] if init_type == "random_c" else None
            plan = initializers.sample_plan(view.shape[1], k[t], norms, rs_plan)
            assert plan.shape == (k[t], int(.2 * view.shape[1]))
            block = np.stack([view[:, plan[c]].mean(axis=1) for c in range(k[t])], axis=1)
            total = total + np.abs(block)
        np.testing.assert_array_equal(total, want[t, t])
    assert rs_ref.rand() == rs_plan.rand()          # the RNG stream was consumed identically


def test_sample_plan_of_a_narrow_relation_is_empty():
    plan = initializers.sample_plan(4, 3, None, np.random.RandomState(0))    # int(0.2 * 4) == 0 -> NaN factors upstream
    assert plan.shape == (3, 0)


# ------------------------------------------------------------------------------------------------ GPU
def _device_init(R, types, n, k, init_type, seed, dtype, storage=None):
    from skfusion.fusion import solver
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G, S = solver.dfmf(R, {}, types, k, max_iter=0, init_type=init_type, random_state=np.random.RandomState(seed),
                           dtype=dtype, storage=storage, device_init=True)
    assert S is None
    return G


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-13), ("float32", 2e-6)])
@pytest.mark.parametrize("init_type", ["random_vcol", "random_c"])
def test_device_init_matches_host_init(init_type, dtype, tol):
    types, n, k, R = _graph(seed=3, sizes=(301, 270, 150), ranks=(12, 9, 20))
    first = {key: mats[0] for key, mats in R.items()}
    want = initializers.initialize(types, n, k, first, init_type, np.random.RandomState(5))
    got = _device_init(R, types, n, k, init_type, 5, dtype)
    for t in types:
        assert rel_fro(want[t, t], got[t, t]) < tol, (t, rel_fro(want[t, t], got[t, t]))


@pytest.mark.gpu
@pytest.mark.parametrize("init_type", ["random_vcol", "random_c"])
def test_device_init_on_the_tensor_core_path(init_type):
    """bf16-stored relations: the selection product runs through the tcgen05 kernels (ones are exact in bf16, fp32 sums)."""
    types, ranks, R = oracle.synthetic_graph(384, n_types=3, rank=64, storage="bfloat16")
    n = oracle.count_objects(R)
    first = {key: mats[0] for key, mats in R.items()}
    want = initializers.initialize(types, n, ranks, first, init_type, np.random.RandomState(2))
    got = _device_init(R, types, n, ranks, init_type, 2, "float32", storage="bfloat16")
    for t in types:
        assert rel_fro(want[t, t], got[t, t]) < 2e-6


@pytest.mark.gpu
def test_device_init_nan_when_fewer_than_five_columns():
    rs = np.random.RandomState(0)
    R = {("a", "b"): [rs.rand(30, 4)]}
    G = _device_init(R, ["a", "b"], {"a": 30, "b": 4}, {"a": 3, "b": 2}, "random_vcol", 0, "float64")
    assert np.isnan(G["a", "a"]).all()             # means over int(0.2 * 4) == 0 columns, as upstream
    assert np.isfinite(G["b", "b"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("algo", ["dfmf", "dfmc"])
def test_fit_seeded_on_the_device_follows_the_oracle(algo):
    from skfusion.fusion import solver
    types, n, k, R = _graph(seed=8, sizes=(120, 90, 75), ranks=(6, 5, 4))
    M = None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if algo == "dfmc":
            rs = np.random.RandomState(1)
            M = {key: [rs.rand(*m.shape) < 0.2 for m in mats] for key, mats in R.items()}
            Go, So = oracle.dfmc(R, M, {}, types, k, max_iter=15, init_type="random_c", random_state=np.random.RandomState(4))
            G, S = solver.dfmc(R, M, {}, types, k, max_iter=15, init_type="random_c", random_state=np.random.RandomState(4),
                               dtype="float64", device_init=True)
        else:
            Go, So = oracle.dfmf(R, {}, types, k, max_iter=15, init_type="random_c", random_state=np.random.RandomState(4))
            G, S = solver.dfmf(R, {}, types, k, max_iter=15, init_type="random_c", random_state=np.random.RandomState(4),
                               dtype="float64", device_init=True)
    for t in types:
        assert rel_fro(Go[t, t], G[t, t]) < 1e-8
    for key in So:
        for l in range(len(So[key])):
            assert rel_fro(So[key][l], S[key][l]) < 1e-7


@pytest.mark.gpu
def test_estimator_uses_device_init_for_device_resident_relations():
    """torch CUDA relations never come back to the host for the default random_c initialisation."""
    import torch
    from skfusion import fusion
    rs = np.random.RandomState(0)
    R12 = rs.rand(300, 280)
    t1, t2 = fusion.ObjectType("T1", 8), fusion.ObjectType("T2", 6)
    host = fusion.Dfmf(max_iter=5, init_type="random_c", random_state=1, dtype="float32", device_init=False).fuse(
        fusion.FusionGraph([fusion.Relation(R12.astype(np.float32).astype(np.float64), t1, t2)]))
    dev = fusion.Dfmf(max_iter=5, init_type="random_c", random_state=1, dtype="float32").fuse(
        fusion.FusionGraph([fusion.Relation(torch.from_numpy(R12.astype(np.float32)).cuda(), t1, t2)]))
    for t in (t1, t2):
        assert rel_fro(host.factor(t), dev.factor(t)) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("init_type", ["random_vcol", "random_c"])
def test_transform_seeded_on_the_device_follows_the_oracle(init_type):
    """DfmfTransform's default seed is the fuser's data-driven one (dfmf.py:169): the target's column means on the GPU."""
    import cases
    from skfusion.fusion import solver
    n, n_new, k = 200, 150, 12
    types, ranks, R = oracle.hashed_graph(n, n_types=3, rank=k, storage="float64")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Go, So = oracle.dfmf(R, {}, types, ranks, max_iter=10, init_type="random", random_state=np.random.RandomState(0))
        tags = {t: cases.Tag(t) for t in types}
        G = {(tags[t], tags[t]): Go[t, t] for t in types}
        S = {(tags[a], tags[b]): So[a, b] for (a, b) in So}
        R_new = {(tags[0], tags[1]): [oracle.hashed_uniform(77, n_new, n)], (tags[0], tags[2]): [oracle.hashed_uniform(78, n_new, n)]}
        rk = {tags[t]: k for t in types}
        for iters in (0, 25):
            ref = oracle.transform(R_new, {}, tags[0], rk, G, S, max_iter=iters, init_type=init_type, random_state=np.random.RandomState(3))
            got = solver.transform(R_new, {}, tags[0], rk, G, S, max_iter=iters, init_type=init_type,
                                   random_state=np.random.RandomState(3), dtype="float64", device_init=True)
            assert rel_fro(ref, got) < 1e-9, (iters, rel_fro(ref, got))


def test_where_the_data_driven_seed_is_computed():
    """options.device_init: 'auto' keeps the reference-exact numpy path for small host graphs."""
    from skfusion.fusion import solver
    from skfusion.fusion.options import resolve, AUTO_FP64_MAX_ENTRIES
    small = {("a", "b"): [np.zeros((10, 12))]}

    class Big(object):                 # shape-only stand-in: the decision never touches the data
        shape = (AUTO_FP64_MAX_ENTRIES // 1000 + 1, 1000)

    assert solver._init_on_device(resolve(n_entries=120), "random_c", small, {}) is False
    assert solver._init_on_device(resolve(n_entries=120), "random", small, {}) is False
    assert solver._init_on_device(resolve(n_entries=120, device_init=True), "random_vcol", small, {}) is True
    assert solver._init_on_device(resolve(n_entries=120, device_init=True), "random", small, {}) is False
    assert solver._init_on_device(resolve(n_entries=None), "random_c", {("a", "b"): [Big()]}, {}) is True
    assert solver._init_on_device(resolve(n_entries=None, device_init=False), "random_c", {("a", "b"): [Big()]}, {}) is False
