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
