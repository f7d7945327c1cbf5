"""skfusion (B200-native): collective matrix tri-factorization behind the scikit-fusion API.

    from skfusion import fusion
    fuser = fusion.Dfmf().fuse(graph)

The decomposition hot path runs in hand-written sm_100a CUDA through include/fz_fusion.h.
"""
from . import fusion  # noqa: F401

__version__ = "0.1.0+b200"
