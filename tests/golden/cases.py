"""Seeded problem definitions shared by make_golden.py (reference side) and the parity tests.

Every case is rebuilt deterministically from numpy RandomState seeds, so the committed fixture only
has to hold the reference's *outputs* (initial factors and trajectory snapshots).  Shapes follow the
reference's own README / unit tests: README 3-type graph (README.md:49-70), parallel relations
(tests/test_multiple_relations.py), rank > n_objects (tests/test_base.py:16-17), masks
(tests/test_dfmc.py:25-39), constraint matrices (datasets/base.py:45-61 dicty ppi in [-0.1, 0]).
"""
import numpy as np


class Tag(object):
    """Minimal stand-in for an object type where identity matters (transform uses ``is``)."""

    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return str(self.name)


def _sparse_sym_constraint(rs, n, density=0.1, scale=0.1):
    th = (rs.rand(n, n) < density) * (rs.rand(n, n) - 0.8) * scale
    return (th + th.T) / 2


def fit_cases():
    c = {}

    rs = np.random.RandomState(100)
    c["readme3"] = dict(
        algo="dfmf", types=["t1", "t2", "t3"], ranks={"t1": 10, "t2": 20, "t3": 30},
        R={("t1", "t2"): [rs.rand(50, 100)], ("t1", "t3"): [rs.rand(50, 40)], ("t2", "t3"): [rs.rand(100, 40)]},
        Theta={}, M=None, init_type="random_c", seed=0, max_iter=50, snapshots=[0, 9, 49])

    rs = np.random.RandomState(101)
    c["multi_theta"] = dict(
        algo="dfmf", types=["a", "b", "c"], ranks={"a": 10, "b": 20, "c": 30},
        R={("a", "b"): [rs.rand(50, 100), rs.rand(50, 100) - 0.2], ("a", "c"): [rs.rand(50, 40)],
           ("c", "b"): [rs.rand(40, 100)]},
        Theta={("a", "a"): [_sparse_sym_constraint(rs, 50)], ("b", "b"): [_sparse_sym_constraint(rs, 100),
                                                                      _sparse_sym_constraint(rs, 100)]},
        M=None, init_type="random_vcol", seed=3, max_iter=50, snapshots=[0, 9, 49])

    rs = np.random.RandomState(102)
    c["rank_gt_n"] = dict(
        algo="dfmf", types=["t1", "t2", "t3"], ranks={"t1": 30, "t2": 40, "t3": 40},
        R={("t1", "t2"): [rs.rand(50, 30)], ("t1", "t3"): [rs.rand(50, 40)], ("t2", "t3"): [rs.rand(30, 40)]},
        Theta={}, M=None, init_type="random", seed=1, max_iter=30, snapshots=[0, 9, 29])

    rs = np.random.RandomState(103)
    Rm = {("u", "m"): [rs.rand(60, 80) * 5], ("m", "g"): [(rs.rand(80, 12) < 0.2).astype(float)],
          ("m", "a"): [(rs.rand(80, 70) < 0.1).astype(float), rs.rand(80, 70)]}
    c["completion"] = dict(
        algo="dfmc", types=["u", "m", "g", "a"], ranks={"u": 8, "m": 12, "g": 4, "a": 10},
        R=Rm, Theta={("u", "u"): [_sparse_sym_constraint(rs, 60)]},
        M={("u", "m"): [rs.rand(60, 80) > 0.3], ("m", "g"): [None], ("m", "a"): [None, rs.rand(80, 70) > 0.8]},
        init_type="random_vcol", seed=5, max_iter=40, snapshots=[0, 9, 39])
    return c


def transform_cases():
    c = {}
    rs = np.random.RandomState(200)
    c["project_rows"] = dict(fit="readme3", target="t1",
                             R_new={("t1", "t2"): [rs.rand(10, 100)], ("t1", "t3"): [rs.rand(10, 40)]},
                             Theta={}, init_type="random_c", seed=11, max_iter=60, snapshots=[0, 9, 59])
    rs = np.random.RandomState(201)
    c["project_cols"] = dict(fit="readme3", target="t3",
                             R_new={("t1", "t3"): [rs.rand(50, 9)], ("t2", "t3"): [rs.rand(100, 9)]},
                             Theta={("t3", "t3"): [_sparse_sym_constraint(rs, 9, 0.5, 0.05)]},
                             init_type="random", seed=12, max_iter=60, snapshots=[0, 9, 59])
    return c


def dicty_case(path=None):
    """BASELINE config C2: the dicty graph (Gene 1219 / GO term 116 / Experimental condition 282; ranks
    50 / 15 / 5 as in datasets/base.py:45-61), rebuilt from the derived fixture dicty_matrices.npz."""
    import os
    path = path or os.path.join(os.path.dirname(os.path.abspath(__file__)), "dicty_matrices.npz")
    z = np.load(path)
    shape = tuple(z["ann_shape"])
    ann = np.unpackbits(z["ann_bits"], axis=1)[:, :shape[1]].astype(np.float64)
    expr = z["expr"].astype(np.float64)
    ppi = np.zeros(tuple(z["ppi_shape"]))
    ppi[z["ppi_rows"], z["ppi_cols"]] = z["ppi_vals"].astype(np.float64)
    return dict(algo="dfmf", types=["Gene", "GO term", "Experimental condition"],
                ranks={"Gene": 50, "GO term": 15, "Experimental condition": 5},
                R={("Gene", "GO term"): [ann], ("Gene", "Experimental condition"): [expr]},
                Theta={("Gene", "Gene"): [ppi]}, M=None, init_type="random_vcol", seed=0, max_iter=50)


def movielens_case(path=None):
    """BASELINE config C3 (examples/movielens_completion.py:20-86): users x movies ratings scaled to [0, 1] with
    the unknown (and 10 % hidden) entries masked and mean-filled, movies x genres, movies x actors; ranks
    max(int(0.05 n), 5).  Returned at the function-level seam: R / M / Theta as Dfmc.fuse would marshal them."""
    import os
    path = path or os.path.join(os.path.dirname(os.path.abspath(__file__)), "movielens_matrices.npz")
    z = np.load(path)
    n_u, n_m, n_g, n_a = [int(v) for v in z["shape"]]
    R12 = -np.ones((n_u, n_m))
    R12[z["r_u"], z["r_m"]] = z["r_v"].astype(np.float64)
    unknown = R12 < 0
    known = R12[~unknown]
    R12 = (R12 - known.min()) / (known.max() - known.min())
    hide = np.logical_and(np.random.RandomState(0).random_sample(R12.shape) > 0.9, ~unknown)
    mask = np.logical_or(unknown, hide)
    filled = R12.copy()
    filled[mask] = R12[~mask].mean()                      # Relation.filled() with fill_value='mean'
    R23 = np.zeros((n_m, n_g))
    R23[z["g_m"], z["g_g"]] = 1.
    R24 = np.zeros((n_m, n_a))
    R24[z["a_m"], z["a_a"]] = 1.
    ranks = {"User": max(int(.05 * n_u), 5), "Movie": max(int(.05 * n_m), 5), "Genre": max(int(.05 * n_g), 5),
             "Actor": max(int(.05 * n_a), 5)}
    return dict(algo="dfmc", types=["User", "Movie", "Genre", "Actor"], ranks=ranks,
                R={("User", "Movie"): [filled], ("Movie", "Genre"): [R23], ("Movie", "Actor"): [R24]},
                M={("User", "Movie"): [mask], ("Movie", "Genre"): [None], ("Movie", "Actor"): [None]},
                Theta={}, init_type="random_vcol", seed=0, max_iter=30, truth=R12, hidden=hide)
