"""The C ABI itself: symbols (CPU) and a host-buffer round trip (GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fz_fusion.h")
LIB = os.path.join(ROOT, "scikit-fusion_b200", "libfz_fusion.so")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fz_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_bound_symbols():
    from skfusion import _capi
    assert sorted(_capi.SYMBOLS) == _declared()


def test_header_enums_match_the_binding():
    """dtype / memory / algorithm codes of include/fz_fusion.h are the ones skfusion/_capi.py passes."""
    from skfusion import _capi
    text = open(HEADER).read()
    found = {name: int(val) for name, val in re.findall(r"\b(FZ_[A-Z0-9_]+)\s*=\s*(-?\d+)", text)}
    for name in ("FZ_F64", "FZ_F32", "FZ_BF16", "FZ_U8", "FZ_BF16X3", "FZ_HOST", "FZ_DEVICE", "FZ_DFMF", "FZ_DFMC",
                 "FZ_TERMS_AUTO", "FZ_TERMS_CENTRED1"):
        assert found[name] == getattr(_capi, name), name
    assert _capi.dtype_code("bfloat16x3") == _capi.FZ_BF16X3 == found["FZ_BF16X3"]


def test_storage_of_large_float32_graphs_defaults_to_exact_planes():
    """options.resolve: dtype='auto' keeps float64 for small graphs and takes the float32 engine with the relations as exact bf16
    planes (tensor cores, no bit of R lost) beyond AUTO_FP64_MAX_ENTRIES on one GPU; an explicit choice is never overridden."""
    from skfusion.fusion import options
    big = options.AUTO_FP64_MAX_ENTRIES + 1
    assert options.resolve(n_entries=1000)["dtype"] == "float64" and not options.resolve(n_entries=1000).get("storage")
    got = options.resolve(n_entries=big)
    assert got["dtype"] == "float32" and got["storage"] == "bfloat16x3"
    assert options.resolve(n_entries=big, storage="bfloat16")["storage"] == "bfloat16"
    assert not options.resolve(n_entries=big, dtype="float32").get("storage")        # explicit dtype: exact CUDA-core path
    assert not options.resolve(n_entries=big, n_gpus=2).get("storage")
    with pytest.raises(TypeError):
        options.resolve(n_entries=big, storrage="bfloat16")


def test_library_loads_and_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(LIB)
    for name in _declared():
        assert hasattr(lib, name), "missing export %s" % name
    lib.fz_version.restype = ctypes.c_int
    assert lib.fz_version() >= 100


def test_engine_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without a B200 creating an engine raises with the reason."""
    from conftest import have_gpu
    if have_gpu():
        pytest.skip("a GPU is present")
    from skfusion import _capi
    with pytest.raises(_capi.EngineUnavailable):
        _capi.Engine(0, "float32")


@pytest.mark.gpu
def test_host_buffer_round_trip_and_errors():
    from skfusion import _capi
    rs = np.random.RandomState(0)
    R = rs.rand(70, 33)
    eng = _capi.Engine(0, "float32")
    ti, tj = eng.add_type(70, 6), eng.add_type(33, 5)
    rid = eng.add_relation(ti, tj, R)
    G0i, G0j = rs.rand(70, 6), rs.rand(33, 5)
    eng.set_factor(ti, G0i)
    eng.set_factor(tj, G0j)
    eng.finalize()
    np.testing.assert_allclose(eng.get_factor(ti), G0i.astype(np.float32).astype(np.float64), rtol=0, atol=0)
    with pytest.raises(_capi.EngineError):
        eng.get_backbone(99)
    other = _capi.Engine(0, "float32")
    other.add_type(4, 2)
    with pytest.raises(_capi.EngineError, match="fz_finalize"):      # error text comes from the C side
        other.iterate(_capi.FZ_DFMF, 1)
    with pytest.raises(_capi.EngineError, match="unknown type id"):
        other.add_relation(0, 7, R)
    other.add_type(9, 3)
    with pytest.raises(ValueError, match="object types imply"):      # shapes are checked before the pointer crosses the ABI
        other.add_relation(0, 1, R)
    with pytest.raises(ValueError, match="factor has shape"):
        other.set_factor(0, np.zeros((5, 2)))
    other.close()
    before = eng.launches
    eng.iterate(_capi.FZ_DFMF, 3)
    assert eng.launches > before
    S = eng.get_backbone(rid)
    assert S.shape == (6, 5) and np.isfinite(S).all()
    rec = eng.complete(rid)
    Gi, Gj = eng.get_factor(ti), eng.get_factor(tj)
    np.testing.assert_allclose(rec, Gi @ S @ Gj.T, rtol=2e-4, atol=1e-5)
    eng.close()
