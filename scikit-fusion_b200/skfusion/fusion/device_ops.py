"""Dense products that follow a fit, on the GPU: the completed relation  G_row S G_col^T  (reference
skfusion/fusion/base/base.py:119-167) and the chained profiles  G_row (S_ab S_bc ...) G_col^T  the reference's examples
build from ``chain()`` paths (base.py:69-96, examples/dicty_chaining.py:40-53).  Both are one n_row x n_col rank-k product:
output-bound, run by the engine's tcgen05 kernel with TMA stores (csrc/umma_outer.cuh) in the fp32 engine, by the exact
CUDA-core kernel in the fp64 engine (small outputs).  Products below MIN_DEVICE_ENTRIES output entries are not worth a
device round trip and are done with numpy, as upstream does for every size.
"""
import numpy as np

from .. import _capi
from .options import resolve

MIN_DEVICE_ENTRIES = 1 << 20


def gsg(G_row, M, G_col, engine_kwargs=None, out=None):
    """G_row M G_col^T.  ``out`` (optional 2-D torch CUDA tensor, float32 / float64) receives the result on the device and
    is returned; otherwise a float64 numpy array comes back."""
    G_row, G_col, M = np.asarray(G_row), np.asarray(G_col), np.asarray(M, dtype=np.float64)
    n_out = int(G_row.shape[0]) * int(G_col.shape[0])
    if out is None and n_out < MIN_DEVICE_ENTRIES:
        return np.dot(G_row, np.dot(M, G_col.T))
    kwargs = {k: v for k, v in (engine_kwargs or {}).items() if k in ("device", "dtype")}
    opts = resolve(n_entries=n_out, **kwargs)
    device = opts["device"] if out is None else int(out.device.index or 0)
    eng = _capi.Engine(device=device, compute=opts["dtype"])
    try:
        ti = eng.add_type(G_row.shape[0], G_row.shape[1])
        tj = eng.add_type(G_col.shape[0], G_col.shape[1])
        eng.set_factor(ti, G_row)
        eng.set_factor(tj, G_col)
        eng.finalize()
        return eng.profile_product(ti, tj, M, out=out)
    finally:
        eng.close()
