"""Row-sharded DFMF across the GPUs of one box: one process per GPU, torch.distributed for plumbing.

Partition: every object type's rows are split contiguously over the P ranks; relation
R_ij lives as the row block of type i's local rows.  One iteration is

    products   (local)  Gram partials, A_ij = R_ij[p] G_j, B_ij partial = R_ij[p]^T G_i[p], G_i[p]^T A_ij
    all-reduce (fp64)   one packed buffer of every k x k partial (Gram_t and G_i^T R_ij G_j)
    reduce-scatter      each B_ij partial -> the rows of type j this rank owns
    update     (local)  k x k chain (replicated) + multiplicative update of the local rows
    all-gather          the updated factors

In production the collectives run INSIDE the library (``attach_comm``: NCCL on the engine's own stream, the
reduce-scatter of relation r under the products of relation r+1, the Gram all-reduce and the pseudo-inverses under
the first products, the all-gather of type t under the update of type t+1) and the whole loop is one
``engine.iterate`` call.  ``run_iterations`` is the same loop spelled out on the host with torch.distributed
collectives between the engine's phase calls: it drives any ``shard`` object with the five methods below -- the
CUDA engine (``CudaShard``) or a numpy stand-in in the gloo CPU tests of the host logic.
Transform and restarts do not shard: rows / runs are independent, so they run as replicas.
"""
from .. import _capi


class Collectives(object):
    """The three collectives the path needs, on torch tensors, for NCCL (device) or gloo (CPU tests)."""

    def __init__(self, dist, group=None):
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.native_rs = dist.get_backend(group) == "nccl"

    def all_reduce(self, t):
        self.dist.all_reduce(t, group=self.group)

    def reduce_scatter(self, local, full, async_op=False):
        """async_op (NCCL only): returns a work handle; the collective runs on NCCL's stream behind the work already
        enqueued on the current stream, so later kernels of the current stream overlap it."""
        if self.native_rs:
            return self.dist.reduce_scatter_tensor(local, full, group=self.group, async_op=async_op)
        else:  # gloo has no reduce-scatter: all-reduce and keep this rank's chunk
            self.dist.all_reduce(full, group=self.group)
            n = local.numel()
            local.copy_(full.view(-1)[self.rank * n:(self.rank + 1) * n].view_as(local))

    def all_gather(self, full, local):
        if self.native_rs:
            self.dist.all_gather_into_tensor(full, local, group=self.group)
        else:
            chunks = [t.clone() for t in full.view(self.world, -1).unbind(0)]
            self.dist.all_gather(chunks, local.reshape(-1).clone(), group=self.group)
            for i, c in enumerate(chunks):
                full.view(self.world, -1)[i].copy_(c)


def run_iterations(shard, coll, n_iters):
    """The sharded hot loop.  ``shard`` provides products(), update() and the comm buffers.  When the shard can
    run its products relation by relation and the backend is NCCL, the reduce-scatter of relation r is put in
    flight while relation r+1 is streamed."""
    piecewise = coll.world > 1 and getattr(coll, "native_rs", False) and hasattr(shard, "product_relation")
    for _ in range(n_iters):
        if piecewise:
            shard.products_begin()
            pending = []
            for r, (full, local) in enumerate(shard.bpartials()):
                shard.product_relation(r)
                pending.append(coll.reduce_scatter(local, full, async_op=True))
            shard.products_end()
            coll.all_reduce(shard.small())
            for work in pending:
                work.wait()
        else:
            shard.products()
            if coll.world > 1:
                coll.all_reduce(shard.small())
                for full, local in shard.bpartials():
                    coll.reduce_scatter(local, full)
        shard.update()
        if coll.world > 1:
            for full, local in shard.factors():
                coll.all_gather(full, local)


class _RawDevice(object):
    """Expose an engine-owned device buffer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, count, code):
        typestr = {_capi.FZ_F64: "<f8", _capi.FZ_F32: "<f4"}[code]
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class CudaShard(object):
    """One rank's engine in sharded mode plus torch views of its communication buffers."""

    def __init__(self, engine, device, n_relations, n_types, algo=_capi.FZ_DFMF):
        import torch
        self.torch = torch
        self.engine = engine
        self.device = device
        self.algo = algo
        self.n_relations = n_relations
        self.n_types = n_types
        self.world = 1
        self._views = {}

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def _view(self, ptr, count, code):
        key = (ptr, count, code)
        if key not in self._views:
            self._views[key] = self.torch.as_tensor(_RawDevice(ptr, count, code), device=self.device)
        return self._views[key]

    def products(self):
        self.engine.phase_products(self.algo, self._stream())

    def products_begin(self):
        self.engine.phase_products_begin(self.algo, self._stream())

    def product_relation(self, rel):
        self.engine.phase_product_relation(self.algo, rel, self._stream())

    def products_end(self):
        self.engine.phase_products_end(self.algo, self._stream())

    def update(self):
        self.engine.phase_update(self.algo, self._stream())

    def small(self):
        ptr, count = self.engine.comm_small()
        return self._view(ptr, count, _capi.FZ_F64)

    def bpartials(self):
        out = []
        for rel in range(self.n_relations):
            full, local, count, code = self.engine.comm_bpartial(rel)
            out.append((self._view(full, count * self.world, code), self._view(local, count, code)))
        return out

    def factors(self):
        out = []
        for t in range(self.n_types):
            full, count, code = self.engine.comm_factor(t)
            whole = self._view(full, count * self.world, code)
            out.append((whole, whole[self.rank * count:(self.rank + 1) * count]))
        return out


def local_rows(n, world, rank):
    """Row range [lo, hi) of an n-row type owned by ``rank`` (contiguous blocks of ceil(n/world))."""
    m = (n + world - 1) // world
    lo = min(n, rank * m)
    return lo, min(n, lo + m)


def build_sharded_engine(R_local, sizes, ranks, obj_types, G0, world, rank, device, opts):
    """Engine for this rank: global type sizes, local row blocks of the relations, full initial factors."""
    eng = _capi.Engine(device=device, compute=opts.get("dtype", "float32"))
    if opts.get("split_terms") is not None:
        eng.set_split_terms(opts["split_terms"])
    eng.set_shard(world, rank)
    tid = {t: eng.add_type(sizes[t], int(ranks[t])) for t in obj_types}
    rel_ids = {}
    storage = opts.get("storage")
    want = _capi.FZ_BF16 if (storage and _capi.dtype_code(storage) == _capi.FZ_BF16) else eng.compute
    for key, mats in R_local.items():
        rel_ids[key] = []
        for mat in mats:
            # device tensors already in the dtype the engine keeps are used in place; anything else is copied / converted
            borrow = (_capi._is_torch_cuda(mat) and _capi.dtype_code(str(mat.dtype)) == want and mat.stride(1) == 1 and
                      (want != _capi.FZ_BF16 or (mat.stride(0) % 8 == 0 and mat.data_ptr() % 16 == 0)))
            rel_ids[key].append(eng.add_relation(tid[key[0]], tid[key[1]], mat, storage=storage, borrow=borrow))
    for t in obj_types:
        eng.set_factor(tid[t], G0[t, t])
    eng.finalize()
    return eng, tid, rel_ids


def attach_comm(eng, dist, group=None):
    """Give the engine its own NCCL communicator over the ranks of ``dist`` (rank 0 makes the id, torch.distributed
    only carries the 128 bytes).  Afterwards ``eng.iterate`` runs the sharded loop, collectives included."""
    box = [_capi.comm_unique_id() if dist.get_rank(group) == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    eng.comm_init(box[0])


def dfmf_sharded(R_local, obj_types, sizes, obj_type2rank, G0, max_iter, dist, device=0, group=None, collectives="library",
                 **opts):
    """DFMF over the ranks of torch.distributed: every rank passes the row blocks it owns (``local_rows``) and the
    same full initial factors; returns the full factors and the (replicated) backbones on every rank.
    collectives = "library" (NCCL inside the engine; needs the nccl backend) or "host" (torch.distributed calls between
    the engine's phases: the spelled-out loop, also what the gloo tests drive)."""
    coll = Collectives(dist, group)
    opts.pop("n_gpus", None)
    eng, tid, rel_ids = build_sharded_engine(R_local, sizes, obj_type2rank, obj_types, G0, coll.world, coll.rank, device, opts)
    try:
        if collectives == "library" and coll.native_rs and coll.world > 1:
            import torch
            attach_comm(eng, dist, group)
            eng.iterate(_capi.FZ_DFMF, max_iter, torch.cuda.current_stream(device).cuda_stream)
        else:
            shard = CudaShard(eng, device, sum(len(v) for v in rel_ids.values()), len(obj_types))
            shard.world, shard.rank = coll.world, coll.rank
            run_iterations(shard, coll, max_iter)
        G = {(t, t): eng.get_factor(tid[t]) for t in obj_types}
        S = {key: [eng.get_backbone(i) for i in ids] for key, ids in rel_ids.items()}
        return G, S
    finally:
        eng.close()
