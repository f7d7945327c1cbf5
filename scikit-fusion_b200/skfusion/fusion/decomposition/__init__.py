"""Import-path compatibility with the reference layout (skfusion.fusion.decomposition)."""
from ..estimators import Dfmf, Dfmc, DfmfTransform  # noqa: F401
from ..solver import dfmf, dfmc, transform  # noqa: F401
