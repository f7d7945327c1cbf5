"""The mean-centred operand forms of the tensor-core path (include/fz_fusion.h: FZ_TERMS_AUTO / FZ_TERMS_CENTRED1):
the single-term fused kernel (csrc/umma_fused1.cuh) with the fp64 first-order correction of the backbone solve, the
two-term kernel on the centred operand, and the measured gate that chooses between them -- against the float64 oracle fed
the same bf16-rounded relations, at the tolerances stated for the tensor-core path (G <= 1e-3, S <= 5e-3, objective <= 1e-4).
Also: run-to-run reproducibility of the bf16 path, whose B partials are summed by L2 reductions in arrival order."""
import warnings

import numpy as np
import pytest

import cases
import fusion_oracle as oracle
from helpers import rel_fro

pytestmark = pytest.mark.gpu


def _fit(R, Theta, types, ranks, iters, init, seed, **kw):
    from skfusion.fusion import solver
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G, S = solver.dfmf(R, Theta, types, ranks, max_iter=iters, init_type=init, random_state=np.random.RandomState(seed), **kw)
    return G, S, dict(solver.last_fit_info)


def _oracle(R, Theta, types, ranks, iters, init, seed):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return oracle.dfmf(R, Theta, types, ranks, max_iter=iters, init_type=init, random_state=np.random.RandomState(seed))


def _errors(Go, So, G, S, types):
    return (max(rel_fro(Go[t, t], G[t, t]) for t in types), max(rel_fro(So[k][l], S[k][l]) for k in So for l in range(len(So[k]))))


@pytest.mark.parametrize("init", ["random", "random_c", "random_vcol"])
def test_single_term_kernel_meets_the_tolerance_on_the_synthetic_graphs(init):
    n, iters = 1280, 12
    types, ranks, R = oracle.synthetic_graph(n, n_types=3, rank=64, storage="bfloat16")
    Go, So = _oracle(R, {}, types, ranks, iters, init, 0)
    G, S, info = _fit(R, {}, types, ranks, iters, init, 0, dtype="float32", storage="bfloat16", split_terms="centred1", device_init=False)
    eg, es = _errors(Go, So, G, S, types)
    assert info["operand_stats"]["single"] == iters and info["operand_stats"]["two_term"] == 0
    assert eg < 1e-3 and es < 5e-3, (eg, es)
    obj_o, _ = oracle.objective(R, Go, So)
    obj, _ = oracle.objective(R, G, S)
    assert abs(obj - obj_o) / obj_o < 1e-4


def test_first_order_correction_is_what_keeps_the_backbones(monkeypatch):
    """Without B^T lo_j in G_i^T R G_j the single-term backbones on a data-driven seed are off by orders of magnitude
    (scripts/precision_study.py); with it they sit inside the tolerance -- i.e. the correction kernel is live."""
    n, iters = 1280, 6
    types, ranks, R = oracle.synthetic_graph(n, n_types=3, rank=64, storage="bfloat16")
    Go, So = _oracle(R, {}, types, ranks, iters, "random_c", 0)
    G, S, _ = _fit(R, {}, types, ranks, iters, "random_c", 0, dtype="float32", storage="bfloat16", split_terms="centred1", device_init=False)
    assert _errors(Go, So, G, S, types)[1] < 5e-3
    monkeypatch.setenv("FZ_NO_CORR", "1")
    G2, S2, _ = _fit(R, {}, types, ranks, iters, "random_c", 0, dtype="float32", storage="bfloat16", split_terms="centred1", device_init=False)
    assert _errors(Go, So, G2, S2, types)[1] > 3 * _errors(Go, So, G, S, types)[1]


def test_auto_runs_the_two_term_kernel_on_small_ill_conditioned_data():
    """dicty with bf16-rounded relations: a launch-bound graph (the gate is off) on which a single bf16 term is not enough
    (factor error 6e-2 in the precision model): auto keeps two terms and the parity of the two-term path."""
    c = cases.dicty_case()
    Rb = {k: [oracle.bf16_round(m) for m in v] for k, v in c["R"].items()}
    Go, So = _oracle(Rb, c["Theta"], c["types"], c["ranks"], 12, c["init_type"], 0)
    G, S, info = _fit(Rb, c["Theta"], c["types"], c["ranks"], 12, c["init_type"], 0, dtype="float32", storage="bfloat16",
                      split_terms="auto", device_init=False)
    assert info["operand_stats"]["single"] == 0 and info["operand_stats"]["two_term"] == 12
    eg, es = _errors(Go, So, G, S, c["types"])
    assert eg < 1e-3 and es < 5e-3, (eg, es)


@pytest.mark.parametrize("init", ["random", "random_c"])
def test_auto_gate_on_a_bandwidth_bound_graph(init):
    """4 864 objects per type (7e7 relation entries: above the launch-bound limit, so the gate is live): the first iteration
    runs two-term, the measured operand-form error then admits the single-term kernel, and parity holds."""
    n, iters = 4864, 8
    types, ranks, R = oracle.hashed_graph(n, n_types=3, rank=64, storage="bfloat16")
    Go, So = _oracle(R, {}, types, ranks, iters, init, 0)
    G, S, info = _fit(R, {}, types, ranks, iters, init, 0, dtype="float32", storage="bfloat16", split_terms="auto", device_init=False)
    st = info["operand_stats"]
    assert st["two_term"] >= 1 and st["single"] >= iters - 2 and st["single"] + st["two_term"] == iters, st
    assert 0.0 < st["err"] < 1e-4 and st["cond"] > 1.0
    eg, es = _errors(Go, So, G, S, types)
    assert eg < 1e-3 and es < 5e-3, (eg, es)


@pytest.mark.parametrize("terms", [2, "centred1"])
def test_run_to_run_reproducibility_of_the_tensor_core_path(terms):
    """The B partials are summed by fp32 reductions in L2 in arrival order (cp.reduce.async.bulk ... add), A likewise when a
    row group is shared by CTAs: two identical fits may differ in the last bits of every sum.  Bound it: after 10 iterations
    the factors of two runs agree to 1e-5 -- two orders below the parity tolerance -- and the fp64 k x k chain keeps the
    backbones within 1e-4."""
    n, iters = 2048, 10
    types, ranks, R = oracle.hashed_graph(n, n_types=3, rank=64, storage="bfloat16")
    G1, S1, _ = _fit(R, {}, types, ranks, iters, "random", 0, dtype="float32", storage="bfloat16", split_terms=terms)
    G2, S2, _ = _fit(R, {}, types, ranks, iters, "random", 0, dtype="float32", storage="bfloat16", split_terms=terms)
    dg, ds = _errors(G1, S1, G2, S2, types)
    assert dg < 1e-5 and ds < 1e-4, (dg, ds)
