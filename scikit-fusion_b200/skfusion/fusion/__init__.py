from .graph import FusionGraph, Relation, ObjectType
from .estimators import (FusionBase, FusionFit, FusionTransform, DataFusionError, Dfmf, Dfmc, DfmfTransform)
from .options import engine_options

__all__ = ['FusionGraph', 'Relation', 'ObjectType', 'FusionBase', 'FusionFit', 'FusionTransform',
           'DataFusionError', 'Dfmf', 'Dfmc', 'DfmfTransform', 'engine_options']
