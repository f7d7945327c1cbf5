"""dfmf / dfmc / transform -- the drop-in seam, same signatures as the reference free functions
(skfusion/fusion/decomposition/_dfmf.py:127, _dfmc.py:181, _dfmf.py:330), executed on a B200.

Host work kept here: counting objects, seeding the factors with numpy's RandomState (bit-exact RNG
consumption, initializers.py), early-stopping bookkeeping, logging, the user callback.  Everything
inside ``for iter in range(max_iter)`` runs in the CUDA engine (include/fz_fusion.h); without a
callback / error tracking the whole loop is ONE C call with no host round trips.

Engine-only keyword arguments (all optional, see options.py): device, dtype, storage, split_terms, device_init, n_gpus
(n_gpus > 1: the rows of every object type are sharded over that many GPUs of the box, driven from this process).
Relation matrices may be numpy arrays (copied to the GPU) or torch CUDA tensors (used in place when
their dtype equals the storage dtype).
"""
import logging

import numpy as np

from .. import _capi
from .initializers import initialize, initialize_on_device
from .options import resolve

log = logging.getLogger("skfusion.fusion")

# diagnostics of the most recent fit of this process (engine launches, which fused kernel ran how often, gate measurements)
last_fit_info = {}


def _record(engine, **extra):
    last_fit_info.clear()
    last_fit_info.update(extra)
    try:
        last_fit_info["launches"] = engine.launches
        last_fit_info["operand_stats"] = engine.operand_stats()
    except Exception:       # diagnostics must never fail a fit
        pass


def count_objects(obj_types, R):
    """Objects per type from the relation shapes; the first shape seen wins and mismatches are only
    logged (reference _dfmf.py:95-124)."""
    sizes = {}
    for (row_t, col_t), mats in R.items():
        for mat in mats:
            for axis, obj_type in enumerate((row_t, col_t)):
                seen = sizes.setdefault(obj_type, mat.shape[axis])
                if seen != mat.shape[axis]:
                    log.critical("Relation matrix R_(%s,%s) dimension mismatch" % (row_t, col_t))
    if set(obj_types) != set(sizes.keys()):
        log.critical("Object type specification mismatch")
    return sizes


def _configure_logging(verbose):
    logging.basicConfig(format="%(asctime)s %(levelname)s: %(message)s", datefmt="%m/%d/%Y %I:%M:%S %p",
                        level=50 - verbose)


def _device_of(R, Theta, opts):
    """Device the handle lives on: where the caller's CUDA relations already are (they are used in place), else the
    configured one.  Relations spread over several devices cannot be borrowed by one handle."""
    found = set()
    for block in (R, Theta or {}):
        for mats in block.values():
            for mat in mats:
                if _capi._is_torch_cuda(mat):
                    found.add(int(mat.device.index if mat.device.index is not None else 0))
    if len(found) > 1:
        raise ValueError("relation tensors live on different CUDA devices %s: move them to one device (or use n_gpus)" % sorted(found))
    return found.pop() if found else int(opts["device"])


class _Problem(object):
    """One engine handle plus the id maps between the reference's dict keys and engine ids."""

    def __init__(self, opts, device=None):
        self.opts = opts
        self.engine = _capi.Engine(device=opts["device"] if device is None else device, compute=opts["dtype"])
        if opts.get("split_terms") is not None:
            self.engine.set_split_terms(opts["split_terms"])
        self.type_id = {}
        self.type_order = []
        self.rel_ids = {}          # (ti, tj) -> [relation id per parallel relation]

    def add_types(self, obj_types, sizes, ranks):
        for t in obj_types:
            self.type_id[t] = self.engine.add_type(sizes[t], int(ranks[t]))
            self.type_order.append(t)

    def add_blocks(self, R, Theta, M=None):
        storage = self.opts.get("storage")
        for key, mats in R.items():
            ids = []
            for l, mat in enumerate(mats):
                mask = None
                if M is not None and M.get(key) is not None:
                    mask = M[key][l]
                ids.append(self.engine.add_relation(self.type_id[key[0]], self.type_id[key[1]], mat,
                                                    storage=storage, borrow=self._can_borrow(mat, storage, mask),
                                                    mask=mask))
            self.rel_ids[key] = ids
        # constraint matrices stay exact: in the compute dtype, or (storage bfloat16x3) as exact bf16 planes on the tensor cores
        th_storage = storage if (storage and _capi.dtype_code(storage) == _capi.FZ_BF16X3) else None
        for key, mats in Theta.items():
            for mat in mats:
                self.engine.add_relation(self.type_id[key[0]], self.type_id[key[1]], mat, storage=th_storage)

    def _can_borrow(self, mat, storage, mask):
        if mask is not None or not _capi._is_torch_cuda(mat):
            return False
        # the engine keeps a relation in bf16 only when asked to, otherwise in its compute dtype (fz_add_relation)
        want = _capi.FZ_BF16 if (storage and _capi.dtype_code(storage) == _capi.FZ_BF16) else self.engine.compute
        try:
            have = _capi.dtype_code(str(mat.dtype))
        except ValueError:
            return False
        return have == want and mat.stride(1) == 1 and (want != _capi.FZ_BF16 or (mat.stride(0) % 8 == 0 and mat.data_ptr() % 16 == 0))

    def factors(self):
        return {(t, t): self.engine.get_factor(self.type_id[t]) for t in self.type_order}

    def backbones(self):
        return {key: [self.engine.get_backbone(i) for i in ids] for key, ids in self.rel_ids.items()}

    def n_relations(self):
        return sum(len(v) for v in self.rel_ids.values())

    def relation_index(self, key, l):
        """Position of relation (key, l) in the engine's objective vector (insertion order)."""
        pos = 0
        for k, ids in self.rel_ids.items():
            if k == key:
                return pos + l
            pos += len(ids)
        raise KeyError(key)

    def close(self):
        self.engine.close()


def _fit(algo, R, M, Theta, obj_types, obj_type2rank, max_iter, init_type, stopping, stopping_system, verbose,
         compute_err, callback, random_state, engine_kwargs):
    _configure_logging(verbose)
    opts = resolve(n_entries=_count_entries(R, Theta), **engine_kwargs)
    if int(opts.get("n_gpus") or 1) > 1:
        return _fit_sharded(algo, R, M, Theta, obj_types, obj_type2rank, max_iter, init_type, stopping, stopping_system,
                            compute_err, callback, random_state, opts)
    sizes = count_objects(obj_types, R)
    on_device = _init_on_device(opts, init_type, R, Theta)
    G0 = None
    if not on_device:
        first = {key: _host_view(mats[0]) for key, mats in R.items()} if init_type != "random" else {}
        G0 = initialize(obj_types, sizes, obj_type2rank, first, init_type, random_state)
    if stopping_system:
        compute_err = True

    prob = _Problem(opts, _device_of(R, Theta, opts))
    try:
        prob.add_types(obj_types, sizes, obj_type2rank)
        prob.add_blocks(R, Theta, M)
        if on_device:
            # random_c / random_vcol with the O(k n^2) column means on the GPU; the RNG is consumed on the host
            prob.engine.finalize()
            initialize_on_device(prob.engine, prob.type_id, {key: ids[0] for key, ids in prob.rel_ids.items()}, list(obj_types),
                                 obj_type2rank, list(R.keys()), sizes, init_type, random_state)
            if max_iter <= 0 or stopping or compute_err or callback:
                G0 = prob.factors()
        else:
            for t in obj_types:
                prob.engine.set_factor(prob.type_id[t], G0[t, t])
            prob.engine.finalize()

        interactive = bool(stopping or compute_err or callback)
        if not interactive:
            if max_iter > 0:
                prob.engine.iterate(algo, max_iter)
                return prob.factors(), prob.backbones()
            return G0, None

        err_target, err_system, history = (None, None), (None, None), []
        G, S = G0, None
        for it in range(max_iter):
            if it > 1 and stopping and err_target[1] - err_target[0] < stopping[1]:
                log.info("Early stopping: target matrix change < %5.4f" % stopping[1])
                break
            if it > 1 and stopping_system and err_system[1] - err_system[0] < stopping_system:
                log.info("Early stopping: matrix system change < %5.4f" % stopping_system)
                break
            log.info("Factorization iteration: %d" % it)
            prob.engine.iterate(algo, 1)
            if stopping or compute_err:
                total, per_rel = prob.engine.objective(prob.n_relations())
                if stopping:
                    tkey, tl = _target_of(stopping)
                    err_target = (per_rel[prob.relation_index(tkey, tl)], err_target[0])
                if compute_err:
                    log.info("Error (objective function value): %5.4f" % total)
                    history.append(total)
                    if stopping_system:
                        err_system = (total, err_system[0])
            if callback:
                G, S = prob.factors(), prob.backbones()
                callback(G, S, it)
        if compute_err and history:
            log.info("Violations of optimization objective: %d/%d " % (int(np.sum(np.diff(history) > 0)), len(history)))
        if max_iter > 0:
            G, S = prob.factors(), prob.backbones()
        return G, S
    finally:
        _record(prob.engine, n_gpus=1)
        prob.close()


def _row_block(mat, lo, hi, device):
    """Rows [lo, hi) of a relation for the rank living on ``device``: a numpy view, or the tensor slice moved there."""
    if _capi._is_torch_cuda(mat):
        blk = mat[lo:hi]
        return blk if int(mat.device.index or 0) == device else blk.to("cuda:%d" % device)
    return mat[lo:hi]


def _fit_sharded(algo, R, M, Theta, obj_types, obj_type2rank, max_iter, init_type, stopping, stopping_system, compute_err,
                 callback, random_state, opts):
    """dfmf / dfmc over n_gpus GPUs of this box from ONE process (reference entry: Dfmf.fuse -> dfmf(), dfmf.py:55-106): rank p's
    handle lives on device p and holds the row blocks [lo_p, hi_p) of every relation; the handles form one shard group and the
    library runs one host thread per handle with NCCL for the three exchanges (include/fz_fusion.h: fz_group_*).  Objective,
    early stopping and the callback work as on one GPU; results are read from rank 0 (every rank holds the whole factors)."""
    from .distributed import local_rows
    world = int(opts["n_gpus"])
    base = int(opts["device"])
    sizes = count_objects(obj_types, R)
    on_device = _init_on_device(opts, init_type, R, Theta)
    G0 = None
    if not on_device:
        first = {key: _host_view(mats[0]) for key, mats in R.items()} if init_type != "random" else {}
        G0 = initialize(obj_types, sizes, obj_type2rank, first, init_type, random_state)     # on the host: RNG-exact
    if stopping_system:
        compute_err = True
    probs = []
    try:
        for p in range(world):
            prob = _Problem(opts, base + p)
            probs.append(prob)
            prob.engine.set_shard(world, p)
            prob.add_types(obj_types, sizes, obj_type2rank)
            block = lambda blocks: {key: [_row_block(mat, *local_rows(sizes[key[0]], world, p), base + p) for mat in mats]
                                    for key, mats in blocks.items()}
            masks = None
            if M is not None:       # completion masks shard with their relations' rows (the imputation is row-local)
                masks = {key: [None if m is None else _row_block(m, *local_rows(sizes[key[0]], world, p), base + p) for m in ms]
                         for key, ms in M.items()}
            prob.add_blocks(block(R), block(Theta), masks)
            if not on_device:
                for t in obj_types:
                    prob.engine.set_factor(prob.type_id[t], G0[t, t])
            prob.engine.finalize()
        group = _capi.EngineGroup([prob.engine for prob in probs])
        group.comm_init()
        head = probs[0]
        if on_device:
            # random_c / random_vcol with the O(k n^2) column means on the GPUs: the host keeps the RandomState, every rank
            # computes its rows' share and the group calls carry the collectives (include/fz_fusion.h: fz_group_init_*)
            initialize_on_device(group, head.type_id, {key: ids[0] for key, ids in head.rel_ids.items()}, list(obj_types),
                                 obj_type2rank, list(R.keys()), sizes, init_type, random_state)
            if max_iter <= 0 or stopping or compute_err or callback:
                G0 = head.factors()
        if not (stopping or compute_err or callback):
            if max_iter > 0:
                group.iterate(algo, max_iter)
                return head.factors(), head.backbones()
            return G0, None
        err_target, err_system, history = (None, None), (None, None), []
        G, S = G0, None
        for it in range(max_iter):
            if it > 1 and stopping and err_target[1] - err_target[0] < stopping[1]:
                log.info("Early stopping: target matrix change < %5.4f" % stopping[1])
                break
            if it > 1 and stopping_system and err_system[1] - err_system[0] < stopping_system:
                log.info("Early stopping: matrix system change < %5.4f" % stopping_system)
                break
            log.info("Factorization iteration: %d" % it)
            group.iterate(algo, 1)
            if stopping or compute_err:
                total, per_rel = group.objective(head.n_relations())
                if stopping:
                    tkey, tl = _target_of(stopping)
                    err_target = (per_rel[head.relation_index(tkey, tl)], err_target[0])
                if compute_err:
                    log.info("Error (objective function value): %5.4f" % total)
                    history.append(total)
                    if stopping_system:
                        err_system = (total, err_system[0])
            if callback:
                G, S = head.factors(), head.backbones()
                callback(G, S, it)
        if compute_err and history:
            log.info("Violations of optimization objective: %d/%d " % (int(np.sum(np.diff(history) > 0)), len(history)))
        if max_iter > 0:
            G, S = head.factors(), head.backbones()
        return G, S
    finally:
        if probs:
            _record(probs[0].engine, n_gpus=world)
        for prob in probs:
            prob.close()


def dfmf_runs(R, Theta, obj_types, obj_type2rank, n_run, max_iter=10, init_type="random_vcol", random_state=None, **engine_kwargs):
    """``n_run`` restarts of dfmf on ONE resident copy of the relations, two restarts per pass over them (the reference fans
    restarts out over joblib workers, each with its own copy of the data: decomposition/dfmf.py:87-95).  The restarts draw
    their initial factors from the shared ``random_state`` one after the other, exactly like ``n_run`` sequential dfmf() calls;
    results come back in run order as a list of (G, S).  Needs the tensor-core path (storage='bfloat16', float32 engine) and a
    centred operand form (split_terms 'auto' / 'centred1'); no per-iteration hooks (stopping / callback) -- use dfmf() for those.
    include/fz_fusion.h: fz_pair_iterate."""
    opts = resolve(n_entries=_count_entries(R, Theta), **engine_kwargs)
    if opts.get("split_terms") not in ("auto", "centred1", _capi.FZ_TERMS_AUTO, _capi.FZ_TERMS_CENTRED1):
        raise ValueError("batched restarts need split_terms 'auto' or 'centred1'")
    if int(opts.get("n_gpus") or 1) != 1:
        raise ValueError("batched restarts run on one GPU")
    sizes = count_objects(obj_types, R)
    on_device = _init_on_device(opts, init_type, R, Theta)
    device = _device_of(R, Theta, opts)
    probs = []
    try:
        for run in range(int(n_run)):
            prob = _Problem(opts, device)
            probs.append(prob)
            prob.add_types(obj_types, sizes, obj_type2rank)
            if run == 0:
                prob.add_blocks(R, Theta, None)
            else:       # the other restarts read the first handle's device copies of every matrix
                first, rid = probs[0], 0
                for key, ids in first.rel_ids.items():
                    prob.rel_ids[key] = [prob.engine.add_relation_borrowed(prob.type_id[key[0]], prob.type_id[key[1]],
                                                                           *first.engine.relation_device_ptr(i)) for i in ids]
                    rid += len(ids)
                for key, mats in Theta.items():
                    for _ in mats:
                        prob.engine.add_relation_borrowed(prob.type_id[key[0]], prob.type_id[key[1]],
                                                          *first.engine.relation_device_ptr(rid))
                        rid += 1
            if on_device:
                prob.engine.finalize()
                initialize_on_device(prob.engine, prob.type_id, {key: ids[0] for key, ids in prob.rel_ids.items()}, list(obj_types),
                                     obj_type2rank, list(R.keys()), sizes, init_type, random_state)
            else:
                first_mats = {key: _host_view(mats[0]) for key, mats in R.items()} if init_type != "random" else {}
                G0 = initialize(obj_types, sizes, obj_type2rank, first_mats, init_type, random_state)
                for t in obj_types:
                    prob.engine.set_factor(prob.type_id[t], G0[t, t])
                prob.engine.finalize()
        if max_iter > 0:
            for a in range(0, len(probs) - 1, 2):
                probs[a].engine.pair_iterate(probs[a + 1].engine, max_iter)
            if len(probs) % 2:
                probs[-1].engine.iterate(_capi.FZ_DFMF, max_iter)
        return [(prob.factors(), prob.backbones()) for prob in probs]
    finally:
        if probs:
            _record(probs[0].engine, n_gpus=1, batched_runs=len(probs))
        for prob in reversed(probs):      # the first handle owns the relations the others borrow: it goes last
            prob.close()


def _init_on_device(opts, init_type, R, Theta):
    """Where random_c / random_vcol compute their column means (options.py: device_init)."""
    if init_type == "random":
        return False
    choice = opts.get("device_init", "auto")
    if choice in (True, False):
        return choice
    from .options import AUTO_FP64_MAX_ENTRIES
    resident = any(_capi._is_torch_cuda(mat) for mats in R.values() for mat in mats)
    return resident or _count_entries(R, Theta) > AUTO_FP64_MAX_ENTRIES


def _target_of(stopping):
    """``stopping`` is ((key, l), eps) in dfmc (_dfmc.py:370-374); dfmf's (key, eps) form indexes the
    relation list as an array upstream (_dfmf.py:303-304) -- accepted here as relation 0."""
    target = stopping[0]
    if len(target) == 2 and isinstance(target[0], tuple):
        return target[0], target[1]
    return target, 0


def _count_entries(*blocks):
    total = 0
    for block in blocks:
        for mats in (block or {}).values():
            for mat in mats:
                total += int(mat.shape[0]) * int(mat.shape[1])
    return total


def _host_view(mat):
    if _capi._is_torch_cuda(mat):
        return mat.float().cpu().numpy().astype(np.float64)
    return mat


def dfmf(R, Theta, obj_types, obj_type2rank, max_iter=10, init_type="random_vcol", stopping=None, stopping_system=None,
         verbose=0, compute_err=False, callback=None, random_state=None, n_jobs=1, **engine_kwargs):
    """Data fusion by matrix factorization.  Returns (G, S): G[(t,t)] is n_t x k_t, S[(ti,tj)] a list
    of k_i x k_j backbones, one per parallel relation.  ``n_jobs`` is accepted and ignored (the GPU
    engine replaces the joblib fan-out of _dfmf.py:69-73)."""
    return _fit(_capi.FZ_DFMF, R, None, Theta, obj_types, obj_type2rank, max_iter, init_type, stopping, stopping_system,
                verbose, compute_err, callback, random_state, engine_kwargs)


def dfmc(R, M, Theta, obj_types, obj_type2rank, max_iter=10, init_type="random_vcol", stopping=None, stopping_system=None,
         verbose=0, compute_err=False, callback=None, random_state=None, n_jobs=1, **engine_kwargs):
    """Data fusion by matrix completion: as dfmf, with the entries flagged in M re-imputed from the
    current model every iteration (_dfmc.py:287-292, 319-325).  The caller's R is never modified."""
    return _fit(_capi.FZ_DFMC, R, M, Theta, obj_types, obj_type2rank, max_iter, init_type, stopping, stopping_system,
                verbose, compute_err, callback, random_state, engine_kwargs)


def transform(R_ij, Theta_i, target_obj_type, obj_type2rank, G, S, max_iter=10, init_type="random_c", stopping=None,
              stopping_system=None, verbose=0, compute_err=False, callback=None, random_state=None, **engine_kwargs):
    """Project new objects of ``target_obj_type`` into a fitted latent space: only the target factor
    is updated, the other factors G and all backbones S stay frozen (_dfmf.py:330-458).  Types are
    matched by identity, one backbone per type pair, as upstream."""
    _configure_logging(verbose)
    if not isinstance(random_state, np.random.RandomState):
        random_state = np.random.RandomState(random_state)
    opts = resolve(n_entries=_count_entries(R_ij, Theta_i), **engine_kwargs)
    tgt = target_obj_type
    n_targets = [mats[0].shape[0 if tgt == ti else 1] for (ti, tj), mats in R_ij.items()]
    if len(set(n_targets)) > 1:
        log.critical("Target object type: %s size mismatch" % tgt)
    n_new = n_targets[0]
    on_device = _init_on_device(opts, init_type, R_ij, Theta_i)
    G_i = None
    if not on_device:
        first = {key: _host_view(mats[0]) for key, mats in R_ij.items()} if init_type != "random" else {}
        G_i = initialize([tgt], {tgt: n_new}, obj_type2rank, first, init_type, random_state)[tgt, tgt]
        if max_iter <= 0:
            return G_i

    prob = _Problem(opts, _device_of(R_ij, Theta_i, opts))
    try:
        involved = []
        for (ti, tj), mats in R_ij.items():
            for t in (ti, tj):
                if not any(t is u for u in involved):
                    involved.append(t)
        sizes = {}
        for t in involved:
            sizes[id(t)] = n_new if t is tgt else G[t, t].shape[0]
        for t in involved:
            prob.type_id[id(t)] = prob.engine.add_type(sizes[id(t)], int(obj_type2rank[t]))
        storage = opts.get("storage")
        rel_of = []
        for (ti, tj), mats in R_ij.items():
            for l, mat in enumerate(mats):
                if not (ti is tgt or tj is tgt):
                    continue
                rid = prob.engine.add_relation(prob.type_id[id(ti)], prob.type_id[id(tj)], mat, storage=storage,
                                               borrow=prob._can_borrow(mat, storage, None))
                rel_of.append((rid, (ti, tj), l))
        for key, mats in Theta_i.items():
            for mat in mats:
                prob.engine.add_relation(prob.type_id[id(tgt)], prob.type_id[id(tgt)], mat,
                                         storage=storage if (storage and _capi.dtype_code(storage) == _capi.FZ_BF16X3) else None)
        for t in involved:
            if not (on_device and t is tgt):
                prob.engine.set_factor(prob.type_id[id(t)], G_i if t is tgt else G[t, t])
        prob.engine.finalize()
        if on_device:
            # the target's random_c / random_vcol seed with its column means computed on the GPU (RNG on the host)
            first_rel = {}
            for rid, key, l in rel_of:
                first_rel.setdefault(key, rid)
            n_of = {t: (n_new if t is tgt else G[t, t].shape[0]) for t in involved}
            initialize_on_device(prob.engine, {tgt: prob.type_id[id(tgt)]}, first_rel, [tgt], obj_type2rank,
                                 [key for key in R_ij.keys() if key in first_rel], n_of, init_type, random_state)
            if max_iter <= 0:
                return prob.engine.get_factor(prob.type_id[id(tgt)])
        for rid, key, l in rel_of:
            prob.engine.set_backbone(rid, S[key][l])
        prob.engine.transform_prepare(prob.type_id[id(tgt)])
        if stopping_system:
            compute_err = True
        if callback is None and not compute_err:
            prob.engine.transform_iterate(max_iter)
        else:
            # per-iteration mode: error tracking / early stopping / callback as in _dfmf.py:367-453
            err_system, history = (None, None), []
            for it in range(max_iter):
                if it > 1 and stopping_system and err_system[1] - err_system[0] < stopping_system:
                    log.info("Early stopping: matrix system change < %5.4f" % stopping_system)
                    break
                prob.engine.transform_iterate(1)
                if compute_err:
                    total, _ = prob.engine.objective(len(rel_of))
                    log.info("Error (objective function value): %5.4f" % total)
                    history.append(total)
                    if stopping_system:
                        err_system = (total, err_system[0])
                if callback is not None:
                    callback(prob.engine.get_factor(prob.type_id[id(tgt)]), it)
            if compute_err and history:
                log.info("Violations of optimization objective: %d/%d " % (int(np.sum(np.diff(history) > 0)), len(history)))
        return prob.engine.get_factor(prob.type_id[id(tgt)])
    finally:
        prob.close()
