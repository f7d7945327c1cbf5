"""Edge cases of the domain on the CUDA engine, each against the float64 oracle: degenerate shapes, ranks of 1,
ragged tails around the kernels' tile sizes, zero / constant data, several constraint matrices on one type,
many parallel relations, and relations in both directions between a type pair."""
import warnings

import numpy as np
import pytest

import fusion_oracle as oracle
from helpers import rel_fro

pytestmark = pytest.mark.gpu


def _both(R, Theta, types, ranks, iters=12, dtype="float64", seed=0, **kw):
    from skfusion.fusion import solver
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Go, So = oracle.dfmf(R, Theta, types, ranks, max_iter=iters, init_type="random", random_state=np.random.RandomState(seed))
        G, S = solver.dfmf(R, Theta, types, ranks, max_iter=iters, init_type="random", random_state=np.random.RandomState(seed),
                           dtype=dtype, **kw)
    return Go, So, G, S


def _check(Go, So, G, S, tol=1e-8):
    for key in Go:
        assert rel_fro(Go[key], G[key]) < tol, key
    for key in So:
        for a, b in zip(So[key], S[key]):
            assert rel_fro(a, b) < 10 * tol, key


@pytest.mark.parametrize("shape,ranks", [((1, 1), (1, 1)), ((3, 2), (1, 2)), ((2, 7), (2, 1)), ((65, 33), (3, 5)), ((129, 257), (7, 64))])
def test_degenerate_and_ragged_shapes(shape, ranks):
    rs = np.random.RandomState(1)
    R = {("a", "b"): [rs.rand(*shape) + 0.1]}
    _check(*_both(R, {}, ["a", "b"], {"a": ranks[0], "b": ranks[1]}))


def test_zero_and_constant_relations_stay_finite_and_match():
    R = {("a", "b"): [np.zeros((20, 12))], ("a", "c"): [np.full((20, 9), 0.5)]}
    Go, So, G, S = _both(R, {}, ["a", "b", "c"], {"a": 3, "b": 2, "c": 2}, iters=5)
    for key in G:
        assert np.isfinite(G[key]).all() == np.isfinite(Go[key]).all()
    _check(Go, So, G, S, tol=1e-7)


def test_relations_in_both_directions_and_many_parallel_ones():
    rs = np.random.RandomState(2)
    R = {("a", "b"): [rs.rand(30, 22) for _ in range(4)], ("b", "a"): [rs.rand(22, 30), rs.rand(22, 30)]}
    _check(*_both(R, {}, ["a", "b"], {"a": 4, "b": 6}, iters=15), tol=1e-7)


def test_several_constraints_on_one_type_and_constraints_only_on_others():
    rs = np.random.RandomState(3)

    def sym(n, scale):
        m = (rs.rand(n, n) < 0.2) * (rs.rand(n, n) - 0.7) * scale
        return (m + m.T) / 2
    R = {("a", "b"): [rs.rand(40, 25)], ("b", "c"): [rs.rand(25, 18)]}
    Theta = {("a", "a"): [sym(40, 0.05), sym(40, 0.02), sym(40, 0.03)], ("c", "c"): [sym(18, 0.04)]}
    _check(*_both(R, Theta, ["a", "b", "c"], {"a": 5, "b": 4, "c": 3}, iters=20), tol=1e-7)


@pytest.mark.parametrize("n1,n2,k1,k2", [(127, 129, 64, 64), (256, 255, 33, 64), (390, 130, 64, 17), (64, 1000, 8, 8)])
def test_tensor_core_path_ragged_sizes_and_small_ranks(n1, n2, k1, k2):
    """bf16 storage around the 128 / 256 tile edges and with ranks that are not multiples of 4 (red.global flush)."""
    rs = np.random.RandomState(4)
    R = {("a", "b"): [oracle.bf16_round(rs.rand(n1, n2))]}
    Go, So, G, S = _both(R, {}, ["a", "b"], {"a": k1, "b": k2}, iters=8, dtype="float32", storage="bfloat16", split_terms=2)
    for key in Go:
        assert rel_fro(Go[key], G[key]) < 1e-3, key
    assert rel_fro(So["a", "b"][0], S["a", "b"][0]) < 5e-3


def test_three_split_terms_use_the_two_pass_kernels():
    rs = np.random.RandomState(5)
    R = {("a", "b"): [oracle.bf16_round(rs.rand(300, 200))], ("b", "c"): [oracle.bf16_round(rs.rand(200, 140))]}
    Go, So, G, S = _both(R, {}, ["a", "b", "c"], {"a": 20, "b": 64, "c": 12}, iters=10, dtype="float32", storage="bfloat16",
                         split_terms=3)
    for key in Go:
        assert rel_fro(Go[key], G[key]) < 2e-4, key


def test_max_iter_zero_returns_the_initial_factors():
    from skfusion.fusion import solver
    rs = np.random.RandomState(6)
    R = {("a", "b"): [rs.rand(6, 5)]}
    G, S = solver.dfmf(R, {}, ["a", "b"], {"a": 2, "b": 3}, max_iter=0, init_type="random", random_state=np.random.RandomState(9))
    G0 = np.random.RandomState(9)
    np.testing.assert_array_equal(G["a", "a"], G0.rand(6, 2))
    assert S is None
