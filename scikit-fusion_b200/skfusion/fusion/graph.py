# -*- coding: utf-8 -*-
"""Fusion graph containers: ObjectType, Relation, FusionGraph.

API-compatible with skfusion/fusion/base/fusion_graph.py (reference @ 88dd02c): same class and
method names, same identity rules (object types hash/compare by name, fusion_graph.py:448-452;
relations compare by name-or-uuid but hash by their printed form, :550-567), same insertion-ordered
containers, same fill semantics for unknown values (:464-510).  Graph drawing (:51-172) is out of
scope.  Pure host code: these objects only describe the block structure that
Dfmf/Dfmc/DfmfTransform marshal into the GPU engine.
"""
from collections import OrderedDict
from numbers import Number
from uuid import uuid1

import numpy as np

__all__ = ['FusionGraph', 'Relation', 'ObjectType']


class DataFusionError(Exception):
    pass


class ObjectType(object):
    """A type of objects with its factorization rank (number of latent components)."""

    def __init__(self, name, rank=5):
        self.name = name
        self.rank = rank

    def __str__(self):
        return self.name

    def __repr__(self):
        return '{}("{}")'.format(type(self).__name__, self.name)

    def __hash__(self):
        return hash(str(self))

    def __eq__(self, other):
        return isinstance(other, type(self)) and other.name == self.name

    def __ne__(self, other):
        return not self.__eq__(other)


# ------------------------------------------------------------------------------------------------
# replacing unknown values (masked, NaN, +-inf) before factorization
# ------------------------------------------------------------------------------------------------
def _unknown(x, data_only=False):
    bad = ~np.isfinite(x.data if (data_only and np.ma.is_masked(x)) else x)
    if np.ma.is_masked(x):
        bad = np.logical_or(bad, x.mask)
    return bad


def fill_mean(x):
    """Unknown entries <- mean of the known ones."""
    overall = np.nanmean(x)
    out = x.copy()
    out[_unknown(x)] = overall
    return out


def fill_row(x):
    """Unknown entries <- mean of their row (matrix mean for rows without any known entry)."""
    per_row = np.nanmean(x, 1)
    overall = np.nanmean(x)
    if np.ma.is_masked(x):
        per_row = np.ma.filled(np.ma.masked_invalid(per_row), overall)
        bad = _unknown(x, data_only=True)
    else:
        per_row[np.isnan(per_row)] = overall
        bad = ~np.isfinite(x)
    out = x.copy()
    out[bad] = np.take(per_row, bad.nonzero()[0])
    return out


def fill_col(x):
    return fill_row(x.T).T


def fill_const(x, const):
    out = x.copy()
    out[~np.isfinite(x)] = const
    if np.ma.is_masked(x):
        out.data[x.mask] = const
    return out


def _is_device_tensor(x):
    return hasattr(x, "data_ptr") and bool(getattr(x, "is_cuda", False))


FILL_CONST = 'const'
FILL_TYPE = {'mean': fill_mean, 'row_mean': fill_row, 'col_mean': fill_col, 'const': fill_const}


class Relation(object):
    """A data matrix relating objects of ``row_type`` (rows) to objects of ``col_type`` (columns).

    fill_value: 'mean' | 'row_mean' | 'col_mean' | number -- how unknown entries are replaced;
    preprocessor / postprocessor: callables applied before fitting / after ``complete``.
    Extra keyword arguments become attributes.
    """

    def __init__(self, data, row_type, col_type, name='', row_names=None, col_names=None, fill_value='mean',
                 row_metadata=None, col_metadata=None, preprocessor=None, postprocessor=None, **kwargs):
        self.data = data
        self.row_type = row_type
        self.col_type = col_type
        self.name = name
        self.row_names = row_names
        self.col_names = col_names
        self.fill_value = fill_value
        self.row_metadata = row_metadata
        self.col_metadata = col_metadata
        self.preprocessor = preprocessor
        self.postprocessor = postprocessor
        for key, value in kwargs.items():
            setattr(self, key, value)
        self._id = name or uuid1()

    def filled(self):
        """A copy of the data with unknown values replaced according to ``fill_value``.  Device-resident data (a torch
        CUDA tensor) is copied and filled on the GPU (fz_fill_unknown); it never visits the host."""
        if _is_device_tensor(self.data):
            import torch
            from .. import _capi
            copy = self.data.clone(memory_format=torch.contiguous_format)   # row-major whatever the caller's strides
            if isinstance(self.fill_value, Number):
                return _capi.fill_unknown(copy, FILL_CONST, self.fill_value)
            return _capi.fill_unknown(copy, self.fill_value)
        if isinstance(self.fill_value, Number):
            return FILL_TYPE[FILL_CONST](self.data, self.fill_value)
        return FILL_TYPE[self.fill_value](self.data)

    def __contains__(self, obj_type):
        return obj_type == self.row_type or obj_type == self.col_type

    def _label(self, show):
        middle = '"%s"' % self.name if self.name else "→"
        return "{}({} {} {})".format(type(self).__name__, show(self.row_type), middle, show(self.col_type))

    def __str__(self):
        return self._label(str)

    def __repr__(self):
        return self._label(repr)

    def __hash__(self):
        return hash(str(self))

    def __eq__(self, other):
        return isinstance(other, type(self)) and other._id == self._id

    def __ne__(self, other):
        return not self.__eq__(other)


class FusionGraph(object):
    """Relations and the object types they connect, in insertion order."""

    def __init__(self, relations=()):
        self.adjacency_matrix = {}          # row_type -> {col_type -> [Relation, ...]}
        self.relations = OrderedDict()
        self.object_types = OrderedDict()
        self._name2relation = {}
        self._name2object_type = {}
        self.add_relations_from(relations)

    # ---- sizes / lookup
    @property
    def n_relations(self):
        return len(self.relations)

    @property
    def n_object_types(self):
        return len(self.object_types)

    def __getitem__(self, key):
        return self.adjacency_matrix.get(key, self._name2relation.get(key, None))

    def __setitem__(self, key, value):
        self.adjacency_matrix[key] = value

    def get_relation(self, name):
        if name not in self._name2relation:
            raise DataFusionError("Relation name unknown")
        return self._name2relation[name]

    def get_object_type(self, name):
        if name not in self._name2object_type:
            raise DataFusionError("Object type name unknown")
        return self._name2object_type[name]

    def get_relations(self, row_type, col_type):
        """Iterator over the (parallel) relations from ``row_type`` to ``col_type``."""
        self._require(row_type, "Object types are not recognized.")
        self._require(col_type, "Object types are not recognized.")
        return iter(self.adjacency_matrix.get(row_type, {}).get(col_type, []))

    def _require(self, object_type, message="Object type not in the fusion graph."):
        if object_type not in self.object_types:
            raise DataFusionError(message)

    # ---- mutation
    def add_relation(self, relation):
        self.relations[relation] = True
        if relation.name:
            self._name2relation[relation.name] = relation
        for ot in (relation.row_type, relation.col_type):
            self.object_types[ot] = True
            self._name2object_type[ot.name] = ot
        row = self.adjacency_matrix.get(relation.row_type, {})
        row[relation.col_type] = row.get(relation.col_type, []) + [relation]
        self.adjacency_matrix[relation.row_type] = row

    def add_relations_from(self, relations):
        for relation in relations:
            self.add_relation(relation)

    def remove_relation(self, relation):
        """Drop a relation; object types left without any relation disappear with it."""
        row, col = relation.row_type, relation.col_type
        self.adjacency_matrix[row][col].remove(relation)
        self.relations.pop(relation)
        if relation.name:
            self._name2relation.pop(relation.name, None)
        if not self.adjacency_matrix[row][col]:
            self.adjacency_matrix[row].pop(col, None)
        if self._isolated(row):
            self.remove_object_type(row)
            if row == col:
                return
        if self._isolated(col):
            self.remove_object_type(col)

    def _isolated(self, object_type):
        return not list(self.in_neighbors(object_type)) and not list(self.out_neighbors(object_type))

    def remove_relations_from(self, relations):
        for relation in relations:
            self.remove_relation(relation)

    def remove_object_type(self, object_type):
        for relation in list(self.relations):
            if object_type in relation and relation in self.relations:
                self.remove_relation(relation)
        self.adjacency_matrix.pop(object_type, None)
        for row in self.adjacency_matrix.values():
            row.pop(object_type, None)
        self._name2object_type.pop(object_type.name, None)
        self.object_types.pop(object_type, None)

    def remove_object_types_from(self, object_types):
        for object_type in object_types:
            self.remove_object_type(object_type)

    # ---- neighbourhood queries
    def out_relations(self, object_type):
        self._require(object_type)
        for rels in self.adjacency_matrix.get(object_type, {}).values():
            for relation in rels:
                yield relation

    def in_relations(self, object_type):
        self._require(object_type)
        for row in self.adjacency_matrix.values():
            for relation in row.get(object_type, []):
                yield relation

    def out_neighbors(self, object_type):
        self._require(object_type)
        return iter(self.adjacency_matrix.get(object_type, {}).keys())

    def in_neighbors(self, object_type):
        self._require(object_type)
        for row_type, row in self.adjacency_matrix.items():
            if len(row.get(object_type, [])) > 0:
                yield row_type

    # ---- names / metadata of the objects of a type, merged over its relations
    def get_names(self, object_type):
        if isinstance(object_type, str):
            object_type = self.get_object_type(object_type)
        size = 0
        for rel in self.out_relations(object_type):
            if rel.row_names:
                return rel.row_names
            size = rel.data.shape[0]
        for rel in self.in_relations(object_type):
            if rel.col_names:
                return rel.col_names
            size = rel.data.shape[1]
        return [str(i) for i in range(size)]

    def get_metadata(self, object_type):
        if isinstance(object_type, str):
            object_type = self.get_object_type(object_type)
        merged = [{} for _ in self.get_names(object_type)]
        for rel in self.out_relations(object_type):
            if rel.row_metadata:
                for dst, src in zip(merged, rel.row_metadata):
                    dst.update(src)
        for rel in self.in_relations(object_type):
            if rel.col_metadata:
                for dst, src in zip(merged, rel.col_metadata):
                    dst.update(src)
        return merged

    def draw_graphviz(self, *args, **kwargs):
        raise NotImplementedError("graph drawing is outside the scope of the B200 engine (SURVEY.md §2)")

    draw_networkx = draw_graphviz

    def __str__(self):
        return "{}(Object types: {}, Relations: {})".format(
            type(self).__name__, len(self.object_types), len(self.relations))

    def __repr__(self):
        return "{}(Object types={}, Relations={})".format(
            type(self).__name__, repr(list(self.object_types.keys())), repr(list(self.relations.keys())))
