"""Unknown-value replacement on device-resident relations: fz_fill_unknown against the host
fill functions, which restate the reference's fill_mean / fill_row / fill_col / fill_const
(skfusion/fusion/base/fusion_graph.py:464-510; pinned by tests/test_fusion_graph.py)."""
import warnings

import numpy as np
import pytest

from skfusion.fusion import graph

pytestmark = pytest.mark.gpu


def _matrix(rows, cols, seed, with_inf=False):
    rs = np.random.RandomState(seed)
    x = rs.rand(rows, cols) * 4 - 1
    x[rs.rand(rows, cols) < 0.15] = np.nan
    x[rows // 2, :] = np.nan                     # a row without any known entry
    x[:, cols // 3] = np.nan                     # a column without any known entry
    if with_inf:
        x[1, 2] = np.inf
    return x


@pytest.mark.parametrize("mode", ["mean", "row_mean", "col_mean", "const"])
@pytest.mark.parametrize("dtype", ["float64", "float32", "bfloat16"])
def test_fill_unknown_matches_host_fill(mode, dtype):
    import torch
    from skfusion import _capi
    x = _matrix(203, 157, 1)
    tdt = {"float64": torch.float64, "float32": torch.float32, "bfloat16": torch.bfloat16}[dtype]
    dev = torch.from_numpy(x).to(tdt).cuda()
    host_in = dev.double().cpu().numpy()          # the values the device sees, in float64
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = graph.fill_const(host_in, 0.25) if mode == "const" else graph.FILL_TYPE[mode](host_in)
    got = _capi.fill_unknown(dev.clone(), mode, 0.25)
    assert got.dtype == tdt
    want_t = torch.from_numpy(want).to(tdt).double().numpy()      # the fill value is rounded to the storage dtype
    tol = {"float64": 1e-14, "float32": 1e-6, "bfloat16": 8e-3}[dtype]
    np.testing.assert_allclose(got.double().cpu().numpy(), want_t, rtol=tol, atol=0)
    known = np.isfinite(host_in)
    np.testing.assert_array_equal(got.double().cpu().numpy()[known], host_in[known])     # known entries untouched


def test_fill_unknown_with_infinite_entries_follows_nanmean():
    import torch
    from skfusion import _capi
    x = _matrix(40, 30, 2, with_inf=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = graph.fill_row(x)
    got = _capi.fill_unknown(torch.from_numpy(x).cuda(), "row_mean").cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-14, equal_nan=True)


def test_relation_filled_stays_on_the_device_and_fits():
    import torch
    from skfusion import fusion
    x = _matrix(120, 90, 3)
    t1, t2 = fusion.ObjectType("A", 6), fusion.ObjectType("B", 5)
    rel_dev = fusion.Relation(torch.from_numpy(x).cuda(), t1, t2, fill_value="col_mean")
    filled = rel_dev.filled()
    assert filled.is_cuda and filled.data_ptr() != rel_dev.data.data_ptr()
    assert bool(torch.isnan(rel_dev.data).any())                  # the caller's tensor is not modified
    rel_host = fusion.Relation(x, t1, t2, fill_value="col_mean")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.testing.assert_allclose(filled.cpu().numpy(), rel_host.filled(), rtol=1e-13)
        a = fusion.Dfmf(max_iter=10, init_type="random", random_state=0, dtype="float64").fuse(fusion.FusionGraph([rel_dev]))
        b = fusion.Dfmf(max_iter=10, init_type="random", random_state=0, dtype="float64").fuse(fusion.FusionGraph([rel_host]))
    np.testing.assert_allclose(a.factor(t1), b.factor(t1), rtol=1e-9)
    np.testing.assert_allclose(a.backbone(rel_dev), b.backbone(rel_host), rtol=1e-8)
