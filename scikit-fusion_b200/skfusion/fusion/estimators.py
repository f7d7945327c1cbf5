"""Estimator classes: Dfmf, Dfmc, DfmfTransform and their accessor bases.

Same public surface as the reference (skfusion/fusion/base/base.py, decomposition/dfmf.py,
decomposition/dfmc.py): constructor keywords, ``fuse`` / ``transform``, ``factor`` / ``backbone`` /
``complete`` / ``chain``, generators over runs when ``n_run > 1``.  The classes only marshal a
FusionGraph into the R / Theta / M block dictionaries and hand them to solver.dfmf / dfmc /
transform, which run on the GPU.  Differences, all deliberate:
  * ``n_jobs`` is accepted and ignored: restarts run one after another on the device, sharing one
    RandomState exactly like the reference does for n_jobs=1 (results there change with n_jobs,
    SURVEY.md F3).
  * extra keywords device / dtype / storage / split_terms / device_init / n_gpus select the engine configuration (options.py).
  * ``complete`` and the new ``chain_profile`` run their n_row x n_col product on the GPU once it is large (device_ops.py).
"""
from collections import defaultdict
from itertools import product

import numpy as np

from . import device_ops, solver
from .. import _capi
from .graph import DataFusionError

__all__ = ['FusionBase', 'FusionFit', 'FusionTransform', 'DataFusionError', 'Dfmf', 'Dfmc', 'DfmfTransform']

_ENGINE_KEYS = ("device", "dtype", "storage", "split_terms", "device_init", "n_gpus", "batch_runs")


class FusionBase(object):
    """Shared accessors.  ``factors_[object_type]`` and ``backbones_[relation]`` are lists over runs."""
    _params = None

    def __init__(self):
        self.factors_ = defaultdict(list)
        self.backbones_ = defaultdict(list)

    def _set_params(self, values):
        values = dict(values)
        values.pop('self', None)
        values.pop('__class__', None)
        engine = values.pop('engine_kwargs', None) or {}
        unknown = set(engine) - set(_ENGINE_KEYS)
        if unknown:
            raise TypeError("unexpected keyword argument(s): %s" % ", ".join(sorted(unknown)))
        self._params = values
        self._engine_kwargs = engine
        self.__dict__.update(values)

    def _run_index(self, run):
        return 0 if run is None else run

    def _per_run(self, getter):
        for run in range(self.n_run):
            yield getter(run)

    def factor(self, object_type, run=None):
        """Latent matrix G of an object type (a generator over runs if n_run > 1 and run is None)."""
        if object_type not in self.fusion_graph.object_types:
            raise DataFusionError("Object type %s is not included in the fusion scheme" % object_type.name)
        if object_type not in self.factors_:
            raise DataFusionError("Unknown object type.")
        if self.n_run > 1 and run is None:
            return self._per_run(lambda r: self.factors_[object_type][r])
        return self.factors_[object_type][self._run_index(run)]

    def chain(self, row_type, col_type):
        """All simple directed paths row_type -> ... -> col_type in the fusion graph, shortest first."""
        frontier = [[row_type]]
        if row_type == col_type:
            yield frontier[0]
        while frontier:
            longer = []
            for path in frontier:
                for nxt in self.fusion_graph.out_neighbors(path[-1]):
                    if nxt in path:
                        continue
                    if nxt == col_type:
                        yield path + [nxt]
                    else:
                        longer.append(path + [nxt])
            frontier = longer

    def chain_profile(self, path, run=None, row_factor=None):
        """The profile the reference's examples compute from one ``chain()`` path (examples/dicty_chaining.py:40-53):
        objects of ``path[0]`` expressed over the objects of ``path[-1]``,  G_first (S_01 S_12 ...) G_last^T  with the
        backbone of the FIRST relation of every hop; a single-type path gives the factor itself.  ``row_factor`` replaces
        the fitted factor of ``path[0]`` (e.g. a DfmfTransform projection of new objects).  The n_first x n_last product
        runs on the GPU (device_ops.gsg)."""
        run = self._run_index(run)
        first = self.factors_[path[0]][run] if row_factor is None else row_factor
        if len(path) == 1:
            return first
        middle = None
        for a, b in zip(path[:-1], path[1:]):
            hop = self.backbones_[next(iter(self.fusion_graph.get_relations(a, b)))][run]
            middle = hop if middle is None else np.dot(middle, hop)
        return device_ops.gsg(first, middle, self.factors_[path[-1]][run], getattr(self, "_engine_kwargs", None))

    def __repr__(self):
        shown = ', '.join('{}={}'.format(k, v) for k, v in self._params.items())
        return '{}({})'.format(type(self).__name__, shown)

    __str__ = __repr__


class FusionFit(FusionBase):
    """Accessors of a fitted model: backbones and completed relations."""

    def __init__(self):
        super(FusionFit, self).__init__()

    def _check_relation_types(self, relation, message):
        types = self.fusion_graph.object_types
        if relation.row_type not in types or relation.col_type not in types:
            raise DataFusionError(message)

    def backbone(self, relation, run=None):
        """Backbone S of a relation (a generator over runs if n_run > 1 and run is None)."""
        self._check_relation_types(relation, 'Object types are not recognized.')
        if relation not in self.backbones_:
            raise DataFusionError("Unknown relation.")
        if self.n_run > 1 and run is None:
            return self._per_run(lambda r: self.backbones_[relation][r])
        return self.backbones_[relation][self._run_index(run)]

    def _reconstruct(self, relation, run):
        G1 = self.factor(relation.row_type, run)
        S12 = self.backbone(relation, run)
        G2 = self.factor(relation.col_type, run)
        approx = device_ops.gsg(G1, S12, G2, getattr(self, "_engine_kwargs", None))
        return relation.postprocessor(approx) if relation.postprocessor else approx

    def complete(self, relation, run=None):
        """Reconstructed relation G_row S G_col^T, post-processed (generator over runs if n_run > 1)."""
        self._check_relation_types(relation, "Object type %s or %s are not included in the fusion scheme" % (
            relation.row_type.name, relation.col_type.name))
        if self.n_run > 1 and run is None:
            return self._per_run(lambda r: self._reconstruct(relation, r))
        return self._reconstruct(relation, self._run_index(run))


class FusionTransform(FusionBase):
    """Accessors of an online projection of new ``target`` objects."""

    def __init__(self):
        super(FusionTransform, self).__init__()

    def _validate_graph(self):
        if self.target not in self.fusion_graph.object_types:
            raise DataFusionError("Object type %s is not included in the fusion scheme." % self.target.name)
        for relation in self.fusion_graph.relations:
            if self.target not in [relation.row_type, relation.col_type]:
                raise DataFusionError("Relation must include target object type: %s." % self.target.name)

    def chain(self, row_type=None, col_type=None):
        if row_type is not None and col_type is not None and row_type is not self.target:
            raise DataFusionError("Starting type should be target type: %s" % self.target.name)
        col_type = row_type if col_type is None else col_type
        return FusionBase.chain(self, self.target, col_type)


def _blocks_for_fit(graph, with_masks):
    """FusionGraph -> (R, Theta, M) exactly as the reference marshals it (dfmf.py:69-85, dfmc.py:69-94):
    pairs in product(object_types, object_types) order, relations in insertion order, data = filled()
    then preprocessor, a surviving numpy mask stripped (and kept as M for completion)."""
    R, T, M = {}, {}, {}
    for row_type, col_type in product(graph.object_types, repeat=2):
        for relation in graph.get_relations(row_type, col_type):
            mask = None
            if with_masks and _capi._is_torch_cuda(relation.data) and relation.row_type != relation.col_type:
                # device-resident relation: its unknown (non-finite) entries are the completion mask, taken on the GPU before
                # the fill.  Like numpy's masked arrays upstream, the mask survives the 'mean' / constant fills and is dropped
                # by 'row_mean' / 'col_mean' (what filled() returns there is a plain array).
                if relation.fill_value not in ("row_mean", "col_mean"):
                    found = _capi.unknown_mask(relation.data)
                    mask = found if bool(found.any().item()) else None
            data = relation.filled()
            if relation.preprocessor:
                data = relation.preprocessor(data)
            if np.ma.is_masked(data):
                mask = data.mask
                data = data.data
            key = (relation.row_type, relation.col_type)
            if relation.row_type != relation.col_type:
                R.setdefault(key, []).append(data)
                M.setdefault(key, []).append(mask)
            else:
                T.setdefault(key, []).append(data)
    return (R, T, M) if with_masks else (R, T, None)


class _Fuser(FusionFit):
    _solver = None
    _uses_masks = False

    def fuse(self, fusion_graph):
        """Fit the collective factorization on ``fusion_graph``; returns self."""
        self.fusion_graph = fusion_graph
        if not isinstance(self.random_state, np.random.RandomState):
            self.random_state = np.random.RandomState(self.random_state)
        # a set, as upstream (dfmf.py:66): iteration order -- hence RNG order -- follows the hashes (F2)
        object_types = set([ot for ot in fusion_graph.object_types])
        ranks = {ot: int(ot.rank) for ot in fusion_graph.object_types}
        R, T, M = _blocks_for_fit(fusion_graph, self._uses_masks)
        common = dict(obj_types=object_types, obj_type2rank=ranks, max_iter=self.max_iter, init_type=self.init_type,
                      stopping=self.stopping, stopping_system=self.stopping_system, verbose=self.verbose,
                      compute_err=self.compute_err, callback=self.callback, random_state=self.random_state,
                      n_jobs=self.n_jobs)
        common.update(self._engine_kwargs)
        self.factors_ = defaultdict(list)
        self.backbones_ = defaultdict(list)

        def keep(G, S):
            for (object_type, _), factor in G.items():
                self.factors_[object_type].append(factor)
            for (row_type, col_type), backbones in (S or {}).items():
                for i, relation in enumerate(fusion_graph.get_relations(row_type, col_type)):
                    self.backbones_[relation].append(backbones[i])

        if self._batch_restarts(R, T):
            engine_kwargs = {k: v for k, v in self._engine_kwargs.items() if k != "batch_runs"}
            for G, S in solver.dfmf_runs(R, T, object_types, ranks, self.n_run, max_iter=self.max_iter, init_type=self.init_type,
                                         random_state=self.random_state, **engine_kwargs):
                keep(G, S)
            return self
        common.pop("batch_runs", None)
        for _ in range(self.n_run):
            if self._uses_masks:
                G, S = solver.dfmc(R=R, M=M, Theta=T, **common)
            else:
                G, S = solver.dfmf(R=R, Theta=T, **common)
            keep(G, S)
        return self

    def _batch_restarts(self, R, T):
        """Whether the n_run restarts go through solver.dfmf_runs (options.py: batch_runs)."""
        from .options import resolve
        opts = resolve(n_entries=solver._count_entries(R, T), **self._engine_kwargs)
        want = opts.get("batch_runs", "auto")
        eligible = (self.n_run > 1 and not self._uses_masks and not (self.stopping or self.stopping_system or self.compute_err or
                                                                       self.callback)
                    and opts["dtype"] == "float32" and opts.get("storage") in ("bfloat16", "bf16")
                    and opts.get("split_terms") in ("auto", "centred1") and int(opts.get("n_gpus") or 1) == 1)
        if want is True and not eligible:
            raise ValueError("batch_runs=True needs n_run > 1, Dfmf, storage='bfloat16' on the float32 engine, split_terms "
                             "'auto' / 'centred1', one GPU and no per-iteration hooks")
        return eligible and want in (True, "auto")


class Dfmf(_Fuser):
    """Data fusion by matrix factorization (Zitnik & Zupan, TPAMI 2014) on the B200 engine."""
    _uses_masks = False

    def __init__(self, max_iter=100, init_type='random_c', n_run=1, stopping=None, stopping_system=None, verbose=0,
                 compute_err=False, callback=None, random_state=None, n_jobs=1, **engine_kwargs):
        super(Dfmf, self).__init__()
        self._set_params(vars())


class Dfmc(_Fuser):
    """Data fusion by matrix completion: masked entries are re-imputed from the model every iteration."""
    _uses_masks = True

    def __init__(self, max_iter=100, init_type='random_c', n_run=1, stopping=None, stopping_system=None, verbose=0,
                 compute_err=False, callback=None, random_state=None, n_jobs=1, **engine_kwargs):
        super(Dfmc, self).__init__()
        self._set_params(vars())


class DfmfTransform(FusionTransform):
    """Online projection of new objects of one type into the latent space of a fitted fuser."""

    def __init__(self, max_iter=100, init_type=None, n_run=1, stopping=None, stopping_system=None, fill_value=0,
                 verbose=0, compute_err=False, callback=None, random_state=None, n_jobs=1, **engine_kwargs):
        super(DfmfTransform, self).__init__()
        self._set_params(vars())

    def transform(self, target, fusion_graph, fuser):
        self.target = target
        self.fusion_graph = fusion_graph
        self.fuser = fuser
        self._validate_graph()
        init_type = self.init_type if self.init_type is not None else fuser.init_type
        if not isinstance(self.random_state, np.random.RandomState):
            self.random_state = np.random.RandomState(self.random_state)
        ranks = {ot: int(ot.rank) for ot in fusion_graph.object_types}

        R, T = {}, {}
        for row_type, col_type in product(fusion_graph.object_types, repeat=2):
            for relation in fusion_graph.get_relations(row_type, col_type):
                data = relation.preprocessor(relation.data) if relation.preprocessor else relation.data
                if _capi._is_torch_cuda(data):
                    _capi.fill_unknown(data, "const", self.fill_value)      # in place, on the device
                else:
                    if np.ma.is_masked(data):
                        data.fill_value = self.fill_value
                        data = data.filled()
                    data[~np.isfinite(data)] = self.fill_value       # in place, as upstream (dfmf.py:185)
                block = R if relation.row_type != relation.col_type else T
                block.setdefault((relation.row_type, relation.col_type), []).append(data)

        self.factors_ = defaultdict(list)
        for run in range(self.n_run):
            G = {(ot, ot): fuser.factor(ot, run) for ot in fuser.fusion_graph.object_types}
            S = {(rel.row_type, rel.col_type): [fuser.backbone(rel, run)]
                 for rel in fuser.fusion_graph.relations if rel.row_type != rel.col_type}
            G_new = solver.transform(R_ij=R, Theta_i=T, target_obj_type=target, obj_type2rank=ranks, G=G, S=S,
                                     max_iter=self.max_iter, init_type=init_type, stopping=self.stopping,
                                     stopping_system=self.stopping_system, verbose=self.verbose,
                                     compute_err=self.compute_err, callback=self.callback,
                                     random_state=self.random_state, **self._engine_kwargs)
            self.factors_[target].append(G_new)
        return self
