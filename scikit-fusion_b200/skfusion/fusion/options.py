"""Engine-side knobs that the reference API has no slot for.

Defaults can be changed process-wide (``engine_options.update(dtype='float64')``), per estimator
(``Dfmf(..., dtype='float64')``) or through the environment:
    SKFUSION_B200_DEVICE, SKFUSION_B200_DTYPE, SKFUSION_B200_STORAGE, SKFUSION_B200_SPLIT_TERMS
  dtype        compute dtype of factors and streamed products: 'float32' (default) or 'float64'
               (parity mode: matches the reference's float64 numpy path to ~1e-12)
  storage      device dtype of relation matrices: None (= dtype), or 'bfloat16' to take the
               tcgen05 tensor-core path (rank <= 64, fp32 engine)
  split_terms  bf16 terms used to represent a factor on the tensor-core path (1..3)
"""
import os

engine_options = {
    "device": int(os.environ.get("SKFUSION_B200_DEVICE", "0")),
    "dtype": os.environ.get("SKFUSION_B200_DTYPE", "float32"),
    "storage": os.environ.get("SKFUSION_B200_STORAGE") or None,
    "split_terms": int(os.environ.get("SKFUSION_B200_SPLIT_TERMS", "2")),
}


def resolve(**overrides):
    opts = dict(engine_options)
    for key, val in overrides.items():
        if val is not None:
            opts[key] = val
    return opts
