"""Initial latent factors (host side).

The draws must consume ``numpy.random.RandomState`` exactly like the reference
(skfusion/fusion/decomposition/_init.py:6-61) or no seed-level parity is possible, so this stays
numpy on the host; the factors are uploaded once before the GPU loop.  Semantics kept:
  random       G_t = rs.rand(n_t, k_t), types visited in ``obj_types`` order            (_init.py:11-17)
  random_vcol  G_t = 1e-5 + sum over relations touching t of |k_t column means|, each mean over the
               first int(0.2*cols) entries of a cumulatively re-shuffled column order     (_init.py:44-61)
  random_c     same, but sampling from the int(0.5*cols) columns of largest 2-norm      (_init.py:20-41)
Only the first relation of a type pair seeds the factors (_dfmf.py:191).  Fewer than 5 columns give
int(0.2*cols) == 0, i.e. means of empty slices -> NaN, as upstream.
"""
import numpy as np

INIT_TYPES = ("random", "random_c", "random_vcol")


def _oriented(pair, matrix, obj_type):
    """The relation seen with ``obj_type`` on the rows."""
    return matrix if obj_type == pair[0] else matrix.T


def _column_pool(view, restrict_to_heavy):
    n_cols = view.shape[1]
    if not restrict_to_heavy:
        return np.arange(n_cols)
    keep = int(.5 * n_cols)
    norms = [np.linalg.norm(view[:, c], 2) for c in range(n_cols)]
    heavy = sorted(enumerate(norms), key=lambda pair: pair[1], reverse=True)[:keep]
    return [col for col, _ in heavy]          # a list: RandomState.shuffle permutes it in place


def _sampled_means(view, pool, rank, random_state):
    sample = int(.2 * view.shape[1])
    block = np.zeros((view.shape[0], rank))
    for col in range(rank):
        random_state.shuffle(pool)
        block[:, col] = view[:, pool[:sample]].mean(axis=1)
    return block


def initialize(obj_types, obj_type2n_obj, obj_type2rank, R, init_type, random_state):
    if init_type not in INIT_TYPES:
        raise KeyError(init_type)
    factors = {}
    for obj_type in obj_types:
        shape = (obj_type2n_obj[obj_type], obj_type2rank[obj_type])
        if init_type == "random":
            factors[obj_type, obj_type] = random_state.rand(*shape)
            continue
        total = 1e-5 * np.ones(shape)
        for pair, matrix in R.items():
            if obj_type not in pair:
                continue
            view = _oriented(pair, matrix, obj_type)
            pool = _column_pool(view, init_type == "random_c")
            total = total + np.abs(_sampled_means(view, pool, shape[1], random_state))
        factors[obj_type, obj_type] = total
    return factors


# ------------------------------------------------------------------------------------------------
# The same initialisation with the O(k n^2) part on the GPU.  The host still owns the RandomState: it draws, in the
# reference's order, the shuffles that pick the sampled columns (so the RNG stream is consumed bit-exactly), and the
# engine computes the column means as a product with a 0/1 selection matrix (include/fz_fusion.h, fz_init_*).
# Differences from the host path are floating-point only: the means are summed in the engine's compute dtype in the
# kernels' order instead of numpy's pairwise order, and random_c ranks columns by norms computed on the device
# (ties between columns whose norms differ in the last bits could order differently).
# ------------------------------------------------------------------------------------------------
def sample_plan(n_cols, rank, norms, random_state):
    """The (rank, p_c) index array the reference's inner loop would use on a relation with ``n_cols`` columns:
    cumulative in-place shuffles of the pool, first p_c entries each time (_init.py:35-38, 55-59).  ``norms`` is None
    for random_vcol, the column 2-norms for random_c."""
    sample = int(.2 * n_cols)
    if norms is None:
        pool = np.arange(n_cols)
    else:
        keep = int(.5 * n_cols)
        heavy = sorted(enumerate(norms), key=lambda pair: pair[1], reverse=True)[:keep]
        pool = np.array([col for col, _ in heavy], dtype=np.int64)   # shuffling an array or a list draws the same numbers
    plan = np.empty((rank, sample), dtype=np.int32)
    for col in range(rank):
        random_state.shuffle(pool)
        plan[col] = pool[:sample]
    return plan


def initialize_on_device(engine, type_id, rel_of_pair, obj_types, obj_type2rank, pairs, n_of, init_type, random_state):
    """Device-side twin of ``initialize`` for init_type random_c / random_vcol.
    rel_of_pair[(ti, tj)] = engine id of the FIRST relation of that pair (_dfmf.py:191); ``pairs`` lists the pairs in
    the R dict's order; n_of[t] = objects of type t."""
    if init_type not in ("random_c", "random_vcol"):
        raise KeyError(init_type)
    for obj_type in obj_types:
        tid = type_id[obj_type]
        engine.init_fill(tid, 1e-5)
        for pair in pairs:
            if obj_type not in pair:
                continue
            rel = rel_of_pair[pair]
            row_role = obj_type == pair[0]
            other = pair[1] if row_role else pair[0]
            norms = engine.relation_norms(rel, 0 if row_role else 1) if init_type == "random_c" else None
            plan = sample_plan(n_of[other], int(obj_type2rank[obj_type]), norms, random_state)
            engine.init_add_sampled_means(tid, rel, plan)
    engine.init_end()
