"""Import-path compatibility with the reference layout (skfusion.fusion.base)."""
from ..estimators import FusionBase, FusionFit, FusionTransform, DataFusionError  # noqa: F401
from ..graph import FusionGraph, Relation, ObjectType  # noqa: F401
