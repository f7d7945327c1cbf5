"""Engine-side knobs that the reference API has no slot for.

Defaults can be changed process-wide (``engine_options.update(dtype='float64')``), per estimator
(``Dfmf(..., dtype='float64')``) or through the environment:
    SKFUSION_B200_DEVICE, SKFUSION_B200_DTYPE, SKFUSION_B200_STORAGE, SKFUSION_B200_SPLIT_TERMS
  dtype        compute dtype of factors and streamed products: 'auto' (default), 'float32' or 'float64'.
               'float64' matches the reference's float64 numpy path to ~1e-8 over a whole fit; 'float32' is the
               fast path (G <= 1e-4..1e-3, S <= 1e-3..5e-3 depending on conditioning).  'auto' takes float64 while
               the graph is small (<= AUTO_FP64_MAX_ENTRIES relation entries: every dataset the reference ships),
               where B200's fp64 rate makes exactness free, and float32 beyond -- or whenever storage is bfloat16.
  storage      device form of the relation matrices: None (= dtype: exact CUDA-core products), 'bfloat16' (the relation
               rounded to bf16, streamed once per iteration through the fused tcgen05 kernels: the fastest path, exact
               for 0/1 data, small integers, ratings ...), or 'bfloat16x3' (float32 engine): the float32 relation kept
               exactly as up to three bf16 planes, one tensor-core pass per non-zero plane -- also for masked relations
               (Dfmc), constraint matrices and ranks above 64.  With dtype='auto' and no storage given, graphs beyond
               AUTO_FP64_MAX_ENTRIES entries on one GPU take 'bfloat16x3' (no bit of R is lost; the factor operand
               carries 16 bits, as on the bfloat16 path).
  split_terms  operand form of the factors on the tensor-core path: 1..3 bf16 terms of the factor itself, or 'auto' --
               the mean-centred form with the single-term / two-term kernel chosen per iteration from a measured error
               estimate (include/fz_fusion.h: FZ_TERMS_AUTO) -- or 'centred1' (always the single-term kernel)
  device_init  where the data-driven initialisations (random_c / random_vcol) compute their column means: 'auto'
               (default: on the GPU when a relation is already device-resident or the graph has more than
               AUTO_FP64_MAX_ENTRIES entries, otherwise with numpy on the host exactly like the reference), True, False.
               The RandomState is consumed identically either way (initializers.py).
  batch_runs   Dfmf(n_run > 1): run the restarts on one resident copy of the relations, two restarts per pass over them
               (solver.dfmf_runs): 'auto' (default: when the fit takes the tensor-core path with a centred operand form and has
               no per-iteration hooks), True (insist), False (one restart after the other, each uploading the data again).
  n_gpus       GPUs of this box a Dfmf fit is spread over (default 1): the rows of every object type are split contiguously
               over devices device .. device + n_gpus - 1, driven from this one process (one host thread per GPU inside the
               library, NCCL over NVLink for the three exchanges).  Dfmc and DfmfTransform stay on
               one GPU (their rows / runs are independent: run replicas).
"""
import os

def _terms(text):
    return text if text in ("auto", "centred1") else int(text)


engine_options = {
    "device": int(os.environ.get("SKFUSION_B200_DEVICE", "0")),
    "dtype": os.environ.get("SKFUSION_B200_DTYPE", "auto"),
    "storage": os.environ.get("SKFUSION_B200_STORAGE") or None,
    "split_terms": _terms(os.environ.get("SKFUSION_B200_SPLIT_TERMS", "auto")),
    "device_init": {"1": True, "0": False}.get(os.environ.get("SKFUSION_B200_DEVICE_INIT", ""), "auto"),
    "n_gpus": int(os.environ.get("SKFUSION_B200_N_GPUS", "1")),
    "batch_runs": {"1": True, "0": False}.get(os.environ.get("SKFUSION_B200_BATCH_RUNS", ""), "auto"),
}


AUTO_FP64_MAX_ENTRIES = 50 * 1000 * 1000


def resolve(n_entries=None, **overrides):
    """Effective engine options; ``n_entries`` (total relation + constraint entries) settles dtype='auto'."""
    opts = dict(engine_options)
    for key, val in overrides.items():
        if key not in engine_options:
            raise TypeError("unknown engine option %r (known: %s)" % (key, ", ".join(sorted(engine_options))))
        if val is not None:
            opts[key] = val
    if opts["dtype"] == "auto":
        small = n_entries is not None and n_entries <= AUTO_FP64_MAX_ENTRIES
        opts["dtype"] = "float64" if (small and not opts.get("storage")) else "float32"
        if (not opts.get("storage") and n_entries is not None and not small and int(opts.get("n_gpus") or 1) == 1
                and os.environ.get("SKFUSION_B200_AUTO_X3", "1") != "0"):
            opts["storage"] = "bfloat16x3"      # large float32 graphs: exact bf16 planes on the tensor cores
    return opts
