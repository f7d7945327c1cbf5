"""complete() and chained profiles on the GPU (csrc/umma_outer.cuh) against numpy float64:
G_row M G_col^T as the reference computes it on the host (skfusion/fusion/base/base.py:119-167,
examples/dicty_chaining.py:40-53)."""
import numpy as np
import pytest

from helpers import rel_fro

pytestmark = pytest.mark.gpu


def _case(n_row, n_col, k_row, k_col, seed=0):
    rs = np.random.RandomState(seed)
    return rs.rand(n_row, k_row), rs.randn(k_row, k_col), rs.rand(n_col, k_col)


@pytest.mark.parametrize("shape", [(128, 128, 64, 64), (1000, 520, 64, 40), (333, 2050, 17, 64), (4096, 3000, 64, 64),
                                   (130, 129, 5, 7)])
def test_tensor_core_profile_product_matches_numpy(shape):
    from skfusion import _capi
    G1, M, G2 = _case(*shape)
    want = G1 @ M @ G2.T
    eng = _capi.Engine(0, "float32")           # straight to the engine: every shape takes the device path
    try:
        ti, tj = eng.add_type(G1.shape[0], G1.shape[1]), eng.add_type(G2.shape[0], G2.shape[1])
        eng.set_factor(ti, G1)
        eng.set_factor(tj, G2)
        eng.finalize()
        got = eng.profile_product(ti, tj, M)
    finally:
        eng.close()
    assert got.shape == want.shape and got.dtype == np.float64
    # two bf16 terms per operand: everything but the 2^-18 cross term, plus fp32 accumulation and fp32 factors
    assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()


def test_profile_product_into_a_device_tensor_and_exact_mode():
    import torch
    from skfusion.fusion import device_ops
    G1, M, G2 = _case(1500, 1100, 48, 64, seed=3)
    want = G1 @ M @ G2.T
    out32 = torch.empty((1500, 1100), dtype=torch.float32, device="cuda")
    res = device_ops.gsg(G1, M, G2, {"dtype": "float32"}, out=out32)
    assert res is out32
    assert np.abs(out32.double().cpu().numpy() - want).max() <= 2e-5 * np.abs(want).max()
    out64 = torch.empty((1500, 1100), dtype=torch.float64, device="cuda")
    device_ops.gsg(G1, M, G2, {"dtype": "float64"}, out=out64)
    assert rel_fro(want, out64.cpu().numpy()) < 1e-13


def test_complete_and_chain_profile_through_the_estimator():
    from skfusion import fusion
    from skfusion.fusion import device_ops
    rs = np.random.RandomState(1)
    a, b, c = fusion.ObjectType("a", 12), fusion.ObjectType("b", 9), fusion.ObjectType("c", 7)
    r_ab = fusion.Relation(rs.rand(1300, 900), a, b)
    r_bc = fusion.Relation(rs.rand(900, 1250), b, c)
    graph = fusion.FusionGraph([r_ab, r_bc])
    fuser = fusion.Dfmf(max_iter=6, init_type="random", random_state=0, dtype="float64").fuse(graph)
    assert 1300 * 900 >= device_ops.MIN_DEVICE_ENTRIES            # this completion runs on the device
    want = fuser.factor(a) @ fuser.backbone(r_ab) @ fuser.factor(b).T
    assert rel_fro(want, fuser.complete(r_ab)) < 1e-12
    paths = list(fuser.chain(a, c))
    assert paths == [[a, b, c]]
    prof = fuser.chain_profile(paths[0])
    want = fuser.factor(a) @ (fuser.backbone(r_ab) @ fuser.backbone(r_bc)) @ fuser.factor(c).T
    assert prof.shape == (1300, 1250) and rel_fro(want, prof) < 1e-12
    assert fuser.chain_profile([a]) is fuser.factor(a)
    new_rows = rs.rand(40, 12)
    assert rel_fro(new_rows @ (fuser.backbone(r_ab) @ fuser.backbone(r_bc)) @ fuser.factor(c).T,
                   fuser.chain_profile(paths[0], row_factor=new_rows)) < 1e-12
