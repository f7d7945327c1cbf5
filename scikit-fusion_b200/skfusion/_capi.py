"""ctypes binding of include/fz_fusion.h (libfz_fusion.so) -- the only door to the GPU.

There is deliberately no CPU fallback: if the library or a B200 is missing, calls raise
``EngineUnavailable`` with the reason.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(_HERE)                      # scikit-fusion_b200/
LIB_PATH = os.path.join(PKG_ROOT, "libfz_fusion.so")
CSRC = os.path.join(PKG_ROOT, "csrc")

FZ_F64, FZ_F32, FZ_BF16, FZ_U8, FZ_BF16X3 = 0, 1, 2, 3, 4
FZ_HOST, FZ_DEVICE = 0, 1
FZ_DFMF, FZ_DFMC = 0, 1
FZ_TERMS_AUTO, FZ_TERMS_CENTRED1 = 0, -1

# every symbol include/fz_fusion.h declares (tests check the library exports all of them)
SYMBOLS = [
    "fz_create", "fz_destroy", "fz_last_error", "fz_version", "fz_launch_count", "fz_set_shard", "fz_comm_unique_id", "fz_comm_init", "fz_group_comm_init", "fz_group_iterate",
    "fz_group_objective", "fz_group_init_fill", "fz_group_relation_norms", "fz_group_init_add_sampled_means",
    "fz_group_init_end", "fz_add_type",
    "fz_add_relation", "fz_set_factor", "fz_set_backbone", "fz_set_split_terms", "fz_operand_stats", "fz_finalize", "fz_iterate", "fz_pair_iterate", "fz_relation_device_ptr",
    "fz_phase_products", "fz_phase_update", "fz_phase_products_begin", "fz_phase_product_relation", "fz_phase_products_end", "fz_comm_small", "fz_comm_bpartial", "fz_comm_factor",
    "fz_transform_prepare", "fz_transform_iterate", "fz_get_factor", "fz_get_backbone", "fz_objective", "fz_complete", "fz_profile_product",
    "fz_fill_uniform", "fz_profile", "fz_profile_read",
    "fz_init_fill", "fz_relation_norms", "fz_init_add_sampled_means", "fz_init_end", "fz_fill_unknown", "fz_unknown_mask",
]


class EngineUnavailable(RuntimeError):
    pass


class EngineError(RuntimeError):
    pass


def nvcc_command(out=LIB_PATH):
    return ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
            "-Xcompiler", "-fPIC", "-o", out, os.path.join(CSRC, "fz_engine.cu")]


def build(force=False, verbose=False):
    """Compile libfz_fusion.so in-tree for sm_100a (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(os.path.dirname(PKG_ROOT), "include", "fz_fusion.h"))
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs if os.path.exists(s))):
        return LIB_PATH
    cmd = nvcc_command()
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise EngineUnavailable("nvcc failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineUnavailable(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    c_void_pp = ctypes.POINTER(ctypes.c_void_p)
    i64, i32, vp = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p
    sig = {
        "fz_create": (i32, [c_void_pp, i32, i32]),
        "fz_destroy": (i32, [vp]),
        "fz_last_error": (ctypes.c_char_p, [vp]),
        "fz_version": (i32, []),
        "fz_launch_count": (i64, [vp]),
        "fz_set_shard": (i32, [vp, i32, i32]),
        "fz_comm_unique_id": (i32, [vp]),
        "fz_comm_init": (i32, [vp, vp]),
        "fz_group_comm_init": (i32, [c_void_pp, i32]),
        "fz_group_iterate": (i32, [c_void_pp, i32, i32, i32]),
        "fz_group_objective": (i32, [c_void_pp, i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
        "fz_group_init_fill": (i32, [c_void_pp, i32, i32, ctypes.c_double]),
        "fz_group_relation_norms": (i32, [c_void_pp, i32, i32, i32, ctypes.POINTER(ctypes.c_double), i64]),
        "fz_group_init_add_sampled_means": (i32, [c_void_pp, i32, i32, i32, ctypes.POINTER(ctypes.c_int32), i32]),
        "fz_group_init_end": (i32, [c_void_pp, i32]),
        "fz_add_type": (i32, [vp, i64, i32]),
        "fz_add_relation": (i32, [vp, i32, i32, vp, i64, i32, i32, i32, i32, vp, i64, i32]),
        "fz_set_factor": (i32, [vp, i32, vp, i64, i32, i32]),
        "fz_set_backbone": (i32, [vp, i32, vp, i64, i32, i32]),
        "fz_set_split_terms": (i32, [vp, i32]),
        "fz_operand_stats": (i32, [vp, ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(ctypes.c_double),
                             ctypes.POINTER(ctypes.c_double)]),
        "fz_finalize": (i32, [vp]),
        "fz_iterate": (i32, [vp, i32, i32, vp]),
        "fz_pair_iterate": (i32, [vp, vp, i32, vp]),
        "fz_relation_device_ptr": (i32, [vp, i32, c_void_pp, ctypes.POINTER(i64), ctypes.POINTER(i32)]),
        "fz_phase_products": (i32, [vp, i32, vp]),
        "fz_phase_update": (i32, [vp, i32, vp]),
        "fz_phase_products_begin": (i32, [vp, i32, vp]),
        "fz_phase_product_relation": (i32, [vp, i32, i32, vp]),
        "fz_phase_products_end": (i32, [vp, i32, vp]),
        "fz_comm_small": (i32, [vp, c_void_pp, ctypes.POINTER(i64)]),
        "fz_comm_bpartial": (i32, [vp, i32, c_void_pp, c_void_pp, ctypes.POINTER(i64), ctypes.POINTER(i32)]),
        "fz_comm_factor": (i32, [vp, i32, c_void_pp, ctypes.POINTER(i64), ctypes.POINTER(i32)]),
        "fz_transform_prepare": (i32, [vp, i32, vp]),
        "fz_transform_iterate": (i32, [vp, i32, vp]),
        "fz_get_factor": (i32, [vp, i32, vp, i64, i32, i32, vp]),
        "fz_get_backbone": (i32, [vp, i32, vp, i64, i32, i32, vp]),
        "fz_objective": (i32, [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), vp]),
        "fz_complete": (i32, [vp, i32, vp, i64, i32, i32, vp]),
        "fz_profile_product": (i32, [vp, i32, i32, vp, i64, i32, i32, vp, i64, i32, i32, vp]),
        "fz_fill_uniform": (i32, [vp, i32, i64, i64, i64, i64, ctypes.c_uint64, vp]),
        "fz_init_fill": (i32, [vp, i32, ctypes.c_double, vp]),
        "fz_relation_norms": (i32, [vp, i32, i32, ctypes.POINTER(ctypes.c_double), vp]),
        "fz_init_add_sampled_means": (i32, [vp, i32, i32, ctypes.POINTER(ctypes.c_int32), i32, vp]),
        "fz_init_end": (i32, [vp]),
        "fz_fill_unknown": (i32, [vp, i32, i64, i64, i64, i32, ctypes.c_double, vp]),
        "fz_unknown_mask": (i32, [vp, i32, i64, i64, i64, vp, i64, vp]),
        "fz_profile": (i32, [vp, i32]),
        "fz_profile_read": (i32, [vp, ctypes.POINTER(i64), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                  ctypes.POINTER(ctypes.c_double)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


_NP2FZ = {np.dtype(np.float64): FZ_F64, np.dtype(np.float32): FZ_F32, np.dtype(np.uint8): FZ_U8,
          np.dtype(np.bool_): FZ_U8}
_DTYPE_NAMES = {"float64": FZ_F64, "float32": FZ_F32, "bfloat16": FZ_BF16, "f64": FZ_F64, "f32": FZ_F32,
                "bf16": FZ_BF16, "bfloat16x3": FZ_BF16X3, "bf16x3": FZ_BF16X3}


def dtype_code(name):
    if isinstance(name, int):
        return name
    try:
        return _DTYPE_NAMES[str(name).replace("torch.", "")]
    except KeyError:
        raise ValueError("unknown dtype %r (float64 | float32 | bfloat16 | bfloat16x3 [storage only])" % (name,))


def comm_unique_id():
    """128-byte NCCL id made by one rank and handed to every rank's Engine.comm_init."""
    L = lib()
    buf = ctypes.create_string_buffer(128)
    rc = L.fz_comm_unique_id(ctypes.cast(buf, ctypes.c_void_p))
    if rc < 0:
        raise EngineError("engine error %d: %s" % (rc, L.fz_last_error(None).decode()))
    return buf.raw


class EngineGroup(object):
    """The handles of one shard group driven from ONE process (rank i = engines[i], each on its own GPU): the group
    calls run one host thread per handle inside the library (include/fz_fusion.h: fz_group_*)."""

    def __init__(self, engines):
        self.engines = list(engines)
        self._L = lib()
        self._arr = (ctypes.c_void_p * len(self.engines))(*[e._h.value for e in self.engines])

    def _ck(self, rc):
        if rc < 0:
            msgs = [self._L.fz_last_error(e._h).decode() for e in self.engines]
            raise EngineError("engine group error %d: %s" % (rc, " | ".join(m for m in msgs if m) or "?"))

    def comm_init(self):
        self._ck(self._L.fz_group_comm_init(self._arr, len(self.engines)))

    def iterate(self, algo, n_iters):
        self._ck(self._L.fz_group_iterate(self._arr, len(self.engines), algo, int(n_iters)))

    def objective(self, n_relations):
        per = (ctypes.c_double * max(1, n_relations))()
        tot = ctypes.c_double()
        self._ck(self._L.fz_group_objective(self._arr, len(self.engines), per, ctypes.byref(tot)))
        return tot.value, list(per)[:n_relations]

    # ---- device-side initialisation on the whole group (same signatures as Engine's, so initializers.initialize_on_device
    # drives either): every call holds a collective and runs on one host thread per handle
    def init_fill(self, t, value):
        self._ck(self._L.fz_group_init_fill(self._arr, len(self.engines), int(t), float(value)))

    def relation_norms(self, rel, axis):
        head = self.engines[0]
        ti, tj = head.rel_types[rel]
        count = head.type_shape[tj][0] if axis == 0 else head.type_shape[ti][0]
        out = np.empty(count, dtype=np.float64)
        self._ck(self._L.fz_group_relation_norms(self._arr, len(self.engines), int(rel), int(axis),
                                                 out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), count))
        return out

    def init_add_sampled_means(self, t, rel, idx):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        self._ck(self._L.fz_group_init_add_sampled_means(self._arr, len(self.engines), int(t), int(rel),
                                                         idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), int(idx.shape[1])))

    def init_end(self):
        self._ck(self._L.fz_group_init_end(self._arr, len(self.engines)))

    def close(self):
        for e in self.engines:
            e.close()


def fill_uniform(tensor, seed, row0=0, stream=0):
    """Fill a 2-D torch CUDA tensor with the engine's counter-based uniform values (see fz_fill_uniform)."""
    if not _is_torch_cuda(tensor) or tensor.dim() != 2 or tensor.stride(1) != 1:
        raise ValueError("fill_uniform needs a 2-D torch CUDA tensor with unit inner stride")
    rc = lib().fz_fill_uniform(ctypes.c_void_p(tensor.data_ptr()), dtype_code(str(tensor.dtype)), int(tensor.stride(0)),
                               int(tensor.shape[0]), int(tensor.shape[1]), int(row0), int(seed), ctypes.c_void_p(stream))
    if rc != 0:
        raise EngineError("fz_fill_uniform failed (%d)" % rc)
    return tensor


FILL_MODES = {"mean": 0, "row_mean": 1, "col_mean": 2, "const": 3}


def fill_unknown(tensor, mode, value=0.0, stream=0):
    """Replace the non-finite entries of a 2-D torch CUDA tensor IN PLACE (see fz_fill_unknown); returns the tensor."""
    if not _is_torch_cuda(tensor) or tensor.dim() != 2 or tensor.stride(1) != 1:
        raise ValueError("fill_unknown needs a 2-D torch CUDA tensor with unit inner stride")
    rc = lib().fz_fill_unknown(ctypes.c_void_p(tensor.data_ptr()), dtype_code(str(tensor.dtype)), int(tensor.stride(0)),
                               int(tensor.shape[0]), int(tensor.shape[1]), FILL_MODES[mode], float(value), ctypes.c_void_p(stream))
    if rc != 0:
        raise EngineError("fz_fill_unknown failed (%d)" % rc)
    return tensor


def unknown_mask(tensor, stream=0):
    """uint8 torch CUDA tensor, 1 where the 2-D torch CUDA tensor holds a non-finite entry (see fz_unknown_mask)."""
    import torch
    if not _is_torch_cuda(tensor) or tensor.dim() != 2 or tensor.stride(1) != 1:
        raise ValueError("unknown_mask needs a 2-D torch CUDA tensor with unit inner stride")
    mask = torch.empty(tuple(tensor.shape), dtype=torch.uint8, device=tensor.device)
    rc = lib().fz_unknown_mask(ctypes.c_void_p(tensor.data_ptr()), dtype_code(str(tensor.dtype)), int(tensor.stride(0)),
                               int(tensor.shape[0]), int(tensor.shape[1]), ctypes.c_void_p(mask.data_ptr()), int(mask.stride(0)),
                               ctypes.c_void_p(stream))
    if rc != 0:
        raise EngineError("fz_unknown_mask failed (%d)" % rc)
    return mask


def _is_torch_cuda(x):
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda") and bool(x.is_cuda)


def _describe(x, allow_mask=False):
    """-> (keepalive, pointer, ld, dtype code, mem) for a 2-D numpy array or torch CUDA tensor."""
    if _is_torch_cuda(x):
        if x.dim() != 2 or x.stride(1) != 1:
            raise ValueError("device matrices must be 2-D with unit inner stride")
        code = dtype_code(str(x.dtype)) if str(x.dtype) != "torch.uint8" else FZ_U8
        if str(x.dtype) == "torch.bool":
            code = FZ_U8
        return x, ctypes.c_void_p(x.data_ptr()), int(x.stride(0)), code, FZ_DEVICE
    if hasattr(x, "data_ptr") and hasattr(x, "is_cuda"):      # torch CPU tensor (e.g. pinned bf16 host buffer)
        if x.dim() != 2 or x.stride(1) != 1:
            raise ValueError("host tensors must be 2-D with unit inner stride")
        name = str(x.dtype)
        code = FZ_U8 if name in ("torch.uint8", "torch.bool") else dtype_code(name)
        return x, ctypes.c_void_p(x.data_ptr()), int(x.stride(0)), code, FZ_HOST
    a = np.asarray(x)
    if a.ndim != 2:
        raise ValueError("expected a 2-D matrix, got shape %r" % (a.shape,))
    if a.dtype not in _NP2FZ:
        a = a.astype(np.float64)
    a = np.ascontiguousarray(a)
    return a, ctypes.c_void_p(a.ctypes.data), int(a.shape[1]) if a.shape[1] else 1, _NP2FZ[a.dtype], FZ_HOST


class Engine(object):
    """Thin RAII wrapper of one fz_engine handle."""

    def __init__(self, device=0, compute="float32"):
        self._L = lib()
        self._h = ctypes.c_void_p()
        self.compute = dtype_code(compute)
        rc = self._L.fz_create(ctypes.byref(self._h), int(device), self.compute)
        if rc != 0:
            msg = self._L.fz_last_error(None).decode()
            self._h = ctypes.c_void_p()
            raise EngineUnavailable("fz_create failed (%d): %s" % (rc, msg))
        self._keep = []
        self.type_shape = []     # (n, k) per type id
        self.rel_types = []      # (ti, tj) per relation id
        self.world, self.rank = 1, 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.fz_destroy(self._h)
            self._h = ctypes.c_void_p()
        self._keep = []

    __del__ = close

    def _ck(self, rc):
        if rc < 0:
            raise EngineError("engine error %d: %s" % (rc, self._L.fz_last_error(self._h).decode()))
        return rc

    # ---- description
    def set_shard(self, world, rank):
        self._ck(self._L.fz_set_shard(self._h, world, rank))
        self.world, self.rank = int(world), int(rank)

    def comm_init(self, unique_id):
        """Join the shard group's NCCL communicator (every rank, same 128-byte id, concurrently); fz_iterate and
        fz_objective then run the collectives inside the library."""
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        self._ck(self._L.fz_comm_init(self._h, ctypes.cast(buf, ctypes.c_void_p)))

    def _local_rows(self, t):
        n = self.type_shape[t][0]
        m = (n + self.world - 1) // self.world
        lo = min(n, self.rank * m)
        return min(n, lo + m) - lo

    def add_type(self, n, k):
        tid = self._ck(self._L.fz_add_type(self._h, int(n), int(k)))
        self.type_shape.append((int(n), int(k)))
        return tid

    def add_relation(self, ti, tj, data, storage=None, borrow=False, mask=None):
        if not (0 <= ti < len(self.type_shape) and 0 <= tj < len(self.type_shape)):
            raise EngineError("unknown type id %r / %r" % (ti, tj))
        want = (self._local_rows(ti), self.type_shape[tj][0])
        if tuple(int(d) for d in data.shape) != want:
            # the C side reads rows x cols straight from the pointer: a wrong shape must never get there
            raise ValueError("relation matrix has shape %r but its object types imply %r" % (tuple(data.shape), want))
        if mask is not None and tuple(int(d) for d in mask.shape) != want:
            raise ValueError("mask has shape %r but the relation is %r" % (tuple(mask.shape), want))
        if not hasattr(data, "data_ptr") and np.asarray(data).dtype in (np.dtype(np.bool_), np.dtype(np.uint8)):
            data = np.asarray(data, dtype=np.float64)   # binary relations are data, not masks (numpy promotes them upstream too)
        elif hasattr(data, "data_ptr") and str(data.dtype) in ("torch.uint8", "torch.bool"):
            data = data.double()
        keep, ptr, ld, code, mem = _describe(data)
        st = code if storage is None else dtype_code(storage)
        mkeep, mptr, mld, mmem = None, None, 0, FZ_HOST
        if mask is not None:
            if not _is_torch_cuda(mask):
                mask = np.ascontiguousarray(np.asarray(mask, dtype=np.uint8))
            mkeep, mptr, mld, _mc, mmem = _describe(mask)
        rid = self._ck(self._L.fz_add_relation(self._h, ti, tj, ptr, ld, code, mem, st, 1 if borrow else 0, mptr, mld, mmem))
        if borrow:
            self._keep.append(keep)
        if st == FZ_BF16 and self.relation_device_ptr(rid)[2] != FZ_BF16:
            import warnings
            warnings.warn("relation %d was asked to be stored in bfloat16 but is kept in the compute dtype (constraint matrices and "
                          "masked relations stay exact: they run on the CUDA-core path, not on the tensor cores)" % rid, RuntimeWarning)
        self.rel_types.append((ti, tj))
        return rid

    def set_factor(self, t, G0):
        if not 0 <= t < len(self.type_shape):
            raise EngineError("unknown type id %r" % (t,))
        if tuple(int(d) for d in G0.shape) != self.type_shape[t]:
            raise ValueError("factor has shape %r, expected %r" % (tuple(G0.shape), self.type_shape[t]))
        keep, ptr, ld, code, mem = _describe(G0)
        self._ck(self._L.fz_set_factor(self._h, t, ptr, ld, code, mem))

    def set_backbone(self, rel, S):
        if not 0 <= rel < len(self.rel_types):
            raise EngineError("unknown relation id %r" % (rel,))
        ti, tj = self.rel_types[rel]
        want = (self.type_shape[ti][1], self.type_shape[tj][1])
        if tuple(int(d) for d in S.shape) != want:
            raise ValueError("backbone has shape %r, expected %r" % (tuple(S.shape), want))
        keep, ptr, ld, code, mem = _describe(S)
        self._ck(self._L.fz_set_backbone(self._h, rel, ptr, ld, code, mem))

    def operand_stats(self):
        """{single, two_term: dfmf iterations run with each fused kernel; paired: of those, batched with a partner handle;
        err, cond: last gate measurement}."""
        a, b, p = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
        e, c = ctypes.c_double(0), ctypes.c_double(0)
        self._ck(self._L.fz_operand_stats(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(p), ctypes.byref(e), ctypes.byref(c)))
        return {"single": a.value, "two_term": b.value, "paired": p.value, "err": e.value, "cond": c.value}

    def set_split_terms(self, terms):
        """1..3 (plain operand form), 'auto' / 0 (centred form, kernel chosen per iteration from a measured error
        estimate) or 'centred1' / -1 (always the single-term kernel); include/fz_fusion.h."""
        terms = {"auto": FZ_TERMS_AUTO, "centred1": FZ_TERMS_CENTRED1}.get(terms, terms)
        self._ck(self._L.fz_set_split_terms(self._h, int(terms)))

    def finalize(self):
        self._ck(self._L.fz_finalize(self._h))

    # ---- loop
    def iterate(self, algo, n_iters, stream=0):
        self._ck(self._L.fz_iterate(self._h, algo, int(n_iters), ctypes.c_void_p(stream)))

    def pair_iterate(self, other, n_iters, stream=0):
        """n_iters dfmf iterations of two restarts at once (this handle = run 0, ``other`` = run 1): fz_pair_iterate."""
        self._ck(self._L.fz_pair_iterate(self._h, other._h, int(n_iters), ctypes.c_void_p(stream)))

    def relation_device_ptr(self, rel):
        p, ld, code = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int()
        self._ck(self._L.fz_relation_device_ptr(self._h, int(rel), ctypes.byref(p), ctypes.byref(ld), ctypes.byref(code)))
        return p.value, ld.value, code.value

    def add_relation_borrowed(self, ti, tj, ptr, ld, code):
        """A relation that lives in another handle's device memory (same device): borrowed, never copied."""
        src = FZ_F32 if int(code) == FZ_BF16X3 else int(code)      # bf16x3: the fp32 master is borrowed, the planes are this handle's
        rid = self._ck(self._L.fz_add_relation(self._h, ti, tj, ctypes.c_void_p(ptr), int(ld), src, FZ_DEVICE, int(code), 1, None, 0,
                                               FZ_HOST))
        self.rel_types.append((ti, tj))
        return rid

    def phase_products(self, algo, stream=0):
        self._ck(self._L.fz_phase_products(self._h, algo, ctypes.c_void_p(stream)))

    def phase_products_begin(self, algo, stream=0):
        self._ck(self._L.fz_phase_products_begin(self._h, algo, ctypes.c_void_p(stream)))

    def phase_product_relation(self, algo, rel, stream=0):
        self._ck(self._L.fz_phase_product_relation(self._h, algo, int(rel), ctypes.c_void_p(stream)))

    def phase_products_end(self, algo, stream=0):
        self._ck(self._L.fz_phase_products_end(self._h, algo, ctypes.c_void_p(stream)))

    def phase_update(self, algo, stream=0):
        self._ck(self._L.fz_phase_update(self._h, algo, ctypes.c_void_p(stream)))

    def transform_prepare(self, target, stream=0):
        self._ck(self._L.fz_transform_prepare(self._h, target, ctypes.c_void_p(stream)))

    def transform_iterate(self, n_iters, stream=0):
        self._ck(self._L.fz_transform_iterate(self._h, int(n_iters), ctypes.c_void_p(stream)))

    def comm_small(self):
        p, n = ctypes.c_void_p(), ctypes.c_int64()
        self._ck(self._L.fz_comm_small(self._h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def comm_bpartial(self, rel):
        f, l, n, d = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int()
        self._ck(self._L.fz_comm_bpartial(self._h, rel, ctypes.byref(f), ctypes.byref(l), ctypes.byref(n), ctypes.byref(d)))
        return f.value, l.value, n.value, d.value

    def comm_factor(self, t):
        f, n, d = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int()
        self._ck(self._L.fz_comm_factor(self._h, t, ctypes.byref(f), ctypes.byref(n), ctypes.byref(d)))
        return f.value, n.value, d.value

    # ---- results (float64 numpy, like the reference returns)
    def get_factor(self, t, stream=0):
        if not 0 <= t < len(self.type_shape):
            raise EngineError("unknown type id %r" % (t,))
        n, k = self.type_shape[t]
        out = np.empty((n, k), dtype=np.float64)
        self._ck(self._L.fz_get_factor(self._h, t, ctypes.c_void_p(out.ctypes.data), k, FZ_F64, FZ_HOST, ctypes.c_void_p(stream)))
        return out

    def get_backbone(self, rel, stream=0):
        if not 0 <= rel < len(self.rel_types):
            raise EngineError("unknown relation id %r" % (rel,))
        ti, tj = self.rel_types[rel]
        ki, kj = self.type_shape[ti][1], self.type_shape[tj][1]
        out = np.empty((ki, kj), dtype=np.float64)
        self._ck(self._L.fz_get_backbone(self._h, rel, ctypes.c_void_p(out.ctypes.data), kj, FZ_F64, FZ_HOST, ctypes.c_void_p(stream)))
        return out

    def objective(self, n_relations, stream=0):
        per = (ctypes.c_double * max(1, n_relations))()
        tot = ctypes.c_double()
        self._ck(self._L.fz_objective(self._h, per, ctypes.byref(tot), ctypes.c_void_p(stream)))
        return tot.value, list(per)[:n_relations]

    def complete(self, rel, stream=0):
        ti, tj = self.rel_types[rel]
        ni, nj = self.type_shape[ti][0], self.type_shape[tj][0]
        out = np.empty((ni, nj), dtype=np.float64)
        self._ck(self._L.fz_complete(self._h, rel, ctypes.c_void_p(out.ctypes.data), nj, FZ_F64, FZ_HOST, ctypes.c_void_p(stream)))
        return out

    def profile_product(self, ti, tj, M, out=None, stream=0):
        """G_ti M G_tj^T (n_ti x n_tj): float64 numpy, or written into ``out`` (a 2-D torch CUDA tensor, float32 / float64)."""
        M = np.ascontiguousarray(np.asarray(M, dtype=np.float64))
        want = (self.type_shape[ti][1], self.type_shape[tj][1])
        if M.shape != want:
            raise ValueError("middle matrix has shape %r, expected %r" % (M.shape, want))
        ni, nj = self.type_shape[ti][0], self.type_shape[tj][0]
        if out is None:
            res = np.empty((ni, nj), dtype=np.float64)
            ptr, ld, code, mem = ctypes.c_void_p(res.ctypes.data), nj, FZ_F64, FZ_HOST
        else:
            if tuple(int(d) for d in out.shape) != (ni, nj):
                raise ValueError("output has shape %r, expected %r" % (tuple(out.shape), (ni, nj)))
            res = out
            _keep, ptr, ld, code, mem = _describe(out)
        self._ck(self._L.fz_profile_product(self._h, int(ti), int(tj), ctypes.c_void_p(M.ctypes.data), M.shape[1], FZ_F64, FZ_HOST,
                                            ptr, ld, code, mem, ctypes.c_void_p(stream)))
        return res

    # ---- factor initialisation on the device (_init.py:20-61; the RNG stays on the host)
    def init_fill(self, t, value, stream=0):
        self._ck(self._L.fz_init_fill(self._h, int(t), float(value), ctypes.c_void_p(stream)))

    def relation_norms(self, rel, axis, stream=0):
        """2-norms of the columns (axis=0) or rows (axis=1) of relation ``rel`` as float64 numpy."""
        ti, tj = self.rel_types[rel]
        count = self.type_shape[tj][0] if axis == 0 else self.type_shape[ti][0]
        out = np.empty(count, dtype=np.float64)
        self._ck(self._L.fz_relation_norms(self._h, int(rel), int(axis), out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                           ctypes.c_void_p(stream)))
        return out

    def init_add_sampled_means(self, t, rel, idx, stream=0):
        """idx: (k_t, p_c) int32 -- per latent column the sampled columns of the relation oriented with t on its rows."""
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        if idx.ndim != 2 or idx.shape[0] != self.type_shape[t][1]:
            raise ValueError("index array has shape %r, expected (%d, p_c)" % (idx.shape, self.type_shape[t][1]))
        self._ck(self._L.fz_init_add_sampled_means(self._h, int(t), int(rel), idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                                   int(idx.shape[1]), ctypes.c_void_p(stream)))

    def init_end(self):
        self._ck(self._L.fz_init_end(self._h))

    def profile(self, enable=True):
        self._ck(self._L.fz_profile(self._h, 1 if enable else 0))

    def profile_read(self):
        """(timed tensor-core launches, summed duration in ms, relation bytes streamed, algorithmic bytes)"""
        n, ms, b, a = ctypes.c_int64(), ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        self._ck(self._L.fz_profile_read(self._h, ctypes.byref(n), ctypes.byref(ms), ctypes.byref(b), ctypes.byref(a)))
        return n.value, ms.value, b.value, a.value

    @property
    def launches(self):
        return int(self._L.fz_launch_count(self._h))
