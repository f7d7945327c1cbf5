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


def test_unknown_mask_of_a_device_relation():
    import torch
    from skfusion import _capi
    x = _matrix(97, 61, 3, with_inf=True)
    for tdt in (torch.float64, torch.float32, torch.bfloat16):
        dev = torch.from_numpy(x).to(tdt).cuda()
        mask = _capi.unknown_mask(dev)
        assert mask.dtype == torch.uint8 and mask.is_cuda
        np.testing.assert_array_equal(mask.cpu().numpy().astype(bool), ~np.isfinite(x))


def test_dfmc_on_a_device_relation_with_unknown_entries_equals_the_masked_host_fit():
    """Dfmc with a torch CUDA relation whose unknown entries are NaN: the mask is extracted on the GPU and the fit equals
    the one of the same data handed over as a numpy masked array (the reference's way, decomposition/dfmc.py:69-94)."""
    import torch
    from skfusion import fusion
    rs = np.random.RandomState(5)
    full = rs.rand(60, 45)
    hidden = rs.rand(60, 45) < 0.2
    a, b = fusion.ObjectType("a", 6), fusion.ObjectType("b", 5)
    host_rel = fusion.Relation(np.ma.masked_array(full, mask=hidden), a, b)
    with_nan = full.copy()
    with_nan[hidden] = np.nan
    dev_rel = fusion.Relation(torch.from_numpy(with_nan).cuda(), a, b)
    kw = dict(max_iter=15, init_type="random", random_state=3, dtype="float64")
    host = fusion.Dfmc(**kw).fuse(fusion.FusionGraph([host_rel]))
    dev = fusion.Dfmc(**kw).fuse(fusion.FusionGraph([dev_rel]))
    assert np.abs(host.factor(a) - dev.factor(a)).max() < 1e-10
    assert np.abs(host.backbone(host_rel) - dev.backbone(dev_rel)).max() < 1e-10
    assert torch.isnan(dev_rel.data).sum().item() == hidden.sum()          # the caller's tensor is untouched


def test_relation_streamed_from_pinned_host_memory_equals_the_resident_one():
    """Out-of-core mode: a bf16 relation borrowed from PINNED host memory is read over PCIe by the same kernels."""
    import torch
    import fusion_oracle as oracle
    from skfusion import _capi
    n, k, iters = 640, 64, 5
    types, ranks, R = oracle.synthetic_graph(n, n_types=2, rank=k, storage="bfloat16")
    mat = torch.from_numpy(R[0, 1][0].astype(np.float32)).to(torch.bfloat16)
    rs = np.random.RandomState(0)
    G0 = [rs.rand(n, k), rs.rand(n, k)]
    out = []
    for where in ("device", "pinned"):
        buf = mat.cuda() if where == "device" else mat.pin_memory()
        eng = _capi.Engine(0, "float32")
        eng.set_split_terms(2)
        t = [eng.add_type(n, k), eng.add_type(n, k)]
        eng.add_relation(t[0], t[1], buf, storage="bfloat16", borrow=True)
        for i in range(2):
            eng.set_factor(t[i], G0[i])
        eng.finalize()
        eng.iterate(_capi.FZ_DFMF, iters)
        out.append((eng.get_factor(t[0]), eng.get_backbone(0)))
        eng.close()
    assert np.abs(out[0][0] - out[1][0]).max() <= 1e-5 * np.abs(out[0][0]).max()
    assert np.abs(out[0][1] - out[1][1]).max() <= 1e-4 * np.abs(out[0][1]).max()
