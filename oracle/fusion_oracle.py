"""CPU oracle for the DFMF / DFMC / transform hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module.  The product path (``scikit-fusion_b200/``) never does; it fails
loudly when its CUDA library is missing.

What this is: a float64 numpy restatement of the reference algorithm (mims-harvard/scikit-fusion
@ 88dd02c), written from the reference's behaviour, one function per reference function, in the
reference's own evaluation order so that results agree to rounding (<= 1e-12 relative, see
tests/test_oracle_pinned.py).  Citations are ``file:line`` under /root/reference/skfusion/fusion/.

Parity status: PINNED.  The reference ships no golden vectors, so the oracle is
pinned (a) in the build container against the real reference imported in memory
(tests/golden/_load_reference.py) and (b) everywhere against fixtures generated from the real
reference by tests/golden/make_golden.py and committed under tests/golden/*.npz.

Conventions shared with the reference (decomposition/_dfmf.py:127-129):
  R[(ti, tj)]      list of 2-D arrays (one per parallel relation), ti != tj
  Theta[(t, t)]    list of square constraint matrices
  M[(ti, tj)]      list of boolean masks or None            (dfmc only)
  G[(t, t)]        n_t x k_t factor;   S[(ti, tj)]  list of k_i x k_j backbones
"""
import numpy as np
import scipy.linalg as spla

EPS64 = np.finfo(float).eps  # clamp of the denominator, _dfmf.py:296


# --------------------------------------------------------------------------- initialisation
def _mean_of_shuffled_columns(Rij, order, rank, take, rs):
    """k columns, each the mean of the first ``take`` entries of a freshly re-shuffled column order.

    ``order`` is shuffled in place and cumulatively, exactly as decomposition/_init.py:34-38,55-60.
    take == 0 gives the mean of an empty slice (NaN + RuntimeWarning), as in the reference.
    """
    out = np.zeros((Rij.shape[0], rank))
    for c in range(rank):
        rs.shuffle(order)
        out[:, c] = Rij[:, order[:take]].mean(axis=1)
    return out


def initialize(obj_types, n_obj, rank, R_first, init_type, rs):
    """decomposition/_init.py:6-61.  R_first maps (ti,tj) -> first relation matrix of that pair."""
    G = {}
    if init_type == "random":                                   # _init.py:11-17
        for t in obj_types:
            G[t, t] = rs.rand(n_obj[t], rank[t])
        return G
    if init_type not in ("random_c", "random_vcol"):
        raise KeyError(init_type)
    for t in obj_types:                                         # _init.py:20-41 / 44-61
        acc = 1e-5 * np.ones((n_obj[t], rank[t]))
        for pair, mat in R_first.items():
            if t not in pair:
                continue
            Rij = mat if t == pair[0] else mat.T
            take = int(.2 * Rij.shape[1])
            if init_type == "random_c":
                keep = int(.5 * Rij.shape[1])
                norms = [np.linalg.norm(Rij[:, c], 2) for c in range(Rij.shape[1])]
                ranked = sorted(enumerate(norms), key=lambda e: e[1], reverse=True)[:keep]
                order = [idx for idx, _ in ranked]              # python list, shuffled in place
            else:
                order = np.arange(Rij.shape[1])
            acc = acc + np.abs(_mean_of_shuffled_columns(Rij, order, rank[t], take, rs))
        G[t, t] = acc
    return G


def count_objects(R):
    """decomposition/_dfmf.py:95-124 (first shape seen wins; mismatches are only logged there)."""
    n = {}
    for (ti, tj), mats in R.items():
        for mat in mats:
            n.setdefault(ti, mat.shape[0])
            n.setdefault(tj, mat.shape[1])
    return n


# --------------------------------------------------------------------------- shared pieces
def _sign_split(x):
    """(positive part, magnitude of negative part) computed as the reference does:
    ``t = x > 0; t*x; (t-1)*x``  (_dfmf.py:256-258) -- kept literal so NaN/inf propagate alike."""
    t = x > 0
    return np.multiply(t, x), np.multiply(t - 1, x)


def _theta_split(Theta):
    """_dfmf.py:203-208: Theta+ = max(Theta,0), Theta- = max(-Theta,0), once."""
    Tp, Tn = {}, {}
    for key, mats in Theta.items():
        for th in mats:
            p, n = _sign_split(th)
            Tp.setdefault(key, []).append(p)
            Tn.setdefault(key, []).append(n)
    return Tp, Tn


def _solve_backbones(R, G):
    """S-update, _dfmf.py:228-239: S_ij = P_i (G_i^T (R_ij (G_j P_j))), P = pinv(nan_to_num(G^T G)),
    with nan_to_num after every block product (_dfmf.py:27,34,40)."""
    P = {key: spla.pinv(np.nan_to_num(np.dot(Gt.T, Gt))) for key, Gt in G.items()}
    GP = {key: np.nan_to_num(np.dot(Gt, P[key])) for key, Gt in G.items()}
    S = {}
    for (ti, tj), mats in R.items():
        if (ti, ti) not in G or (tj, tj) not in G:
            continue
        out = []
        for mat in mats:
            step = np.nan_to_num(np.dot(mat, GP[tj, tj]))
            step = np.nan_to_num(np.dot(G[ti, ti].T, step))
            out.append(np.nan_to_num(np.dot(P[ti, ti], step)))
        S[ti, tj] = out
    return S


def _relation_terms(mat, Gi, Gj, Sij, scrub):
    """Numerator/denominator contributions of one relation to G_i and G_j.

    dfmf (_dfmf.py:249-282) scrubs tmp1/2/4/5 with nan_to_num; dfmc's _update_G_for_Rij
    (_dfmc.py:127-178) and transform (_dfmf.py:394-419) do not."""
    fix = np.nan_to_num if scrub else (lambda a: a)
    t1p, t1n = _sign_split(fix(np.dot(mat, np.dot(Gj, Sij.T))))
    t2p, t2n = _sign_split(fix(np.dot(Sij, np.dot(Gj.T, np.dot(Gj, Sij.T)))))
    t4p, t4n = _sign_split(fix(np.dot(mat.T, np.dot(Gi, Sij))))
    t5p, t5n = _sign_split(fix(np.dot(Sij.T, np.dot(Gi.T, np.dot(Gi, Sij)))))
    return ((t1p + np.dot(Gi, t2n), t1n + np.dot(Gi, t2p)),
            (t4p + np.dot(Gj, t5n), t4n + np.dot(Gj, t5p)))


def _apply_update(G, num, den):
    """_dfmf.py:294-296: G <- G * sqrt(num / max(den, eps64)), all types from the old G."""
    for key in G:
        G[key] = np.multiply(G[key], np.sqrt(np.divide(num[key], np.maximum(den[key], EPS64))))


def objective(R, G, S):
    """Sum over relations of the (un-squared) Frobenius residual, _dfmf.py:306-319."""
    per = []
    for (ti, tj), mats in R.items():
        for l, mat in enumerate(mats):
            approx = np.dot(G[ti, ti], np.dot(S[ti, tj][l], G[tj, tj].T))
            per.append(np.linalg.norm(mat - approx, "fro"))
    return float(sum(per)), per


# --------------------------------------------------------------------------- dfmf
def dfmf(R, Theta, obj_types, obj_type2rank, max_iter=10, init_type="random_vcol", stopping=None,
         stopping_system=None, verbose=0, compute_err=False, callback=None, random_state=None,
         n_jobs=1, G0=None, history=None):
    """decomposition/_dfmf.py:127-327.  ``G0`` (optional) bypasses the initialiser; ``history``
    (optional list) receives the objective per iteration when compute_err is on."""
    n_obj = count_objects(R)
    if G0 is None:
        G = initialize(obj_types, n_obj, obj_type2rank, {k: v[0] for k, v in R.items()}, init_type,
                       random_state)
    else:
        G = {k: np.array(v, dtype=float) for k, v in G0.items()}
    S = None
    if stopping_system:
        compute_err = True
    err_target = (None, None)
    err_system = (None, None)
    Tp, Tn = _theta_split(Theta)

    for it in range(max_iter):
        if it > 1 and stopping and err_target[1] - err_target[0] < stopping[1]:     # :213
            break
        if it > 1 and stopping_system and err_system[1] - err_system[0] < stopping_system:  # :217
            break
        S = _solve_backbones(R, G)
        num = {key: np.zeros(Gt.shape) for key, Gt in G.items()}
        den = {key: np.zeros(Gt.shape) for key, Gt in G.items()}
        for (ti, tj), mats in R.items():
            for l, mat in enumerate(mats):
                (ni, di), (nj, dj) = _relation_terms(mat, G[ti, ti], G[tj, tj], S[ti, tj][l], True)
                num[ti, ti] += ni
                den[ti, ti] += di
                num[tj, tj] += nj
                den[tj, tj] += dj
        for key, mats in Tp.items():                                                 # :285-292
            for th in mats:
                den[key] += np.dot(th, G[key])
        for key, mats in Tn.items():
            for th in mats:
                num[key] += np.dot(th, G[key])
        _apply_update(G, num, den)

        if stopping:
            # The reference indexes R[target]/S[target] as arrays (_dfmf.py:303-304), which only
            # works when they are; dfmc's ((key, l), eps) form is the usable one (_dfmc.py:370-374).
            (key, l), _eps = stopping if isinstance(stopping[0][0], tuple) else ((stopping[0], 0), stopping[1])
            approx = np.dot(G[key[0], key[0]], np.dot(S[key][l], G[key[1], key[1]].T))
            err_target = (np.linalg.norm(R[key][l] - approx), err_target[0])
        if compute_err:
            total, _ = objective(R, G, S)
            if history is not None:
                history.append(total)
            if stopping_system:
                err_system = (total, err_system[0])
        if callback:
            callback(G, S, it)
    return G, S


# --------------------------------------------------------------------------- dfmc
def dfmc(R, M, Theta, obj_types, obj_type2rank, max_iter=10, init_type="random_vcol", stopping=None,
         stopping_system=None, verbose=0, compute_err=False, callback=None, random_state=None,
         n_jobs=1, G0=None, history=None):
    """decomposition/_dfmc.py:181-397: dfmf plus re-imputation of the masked entries."""
    n_obj = count_objects(R)
    if G0 is None:
        G = initialize(obj_types, n_obj, obj_type2rank, {k: v[0] for k, v in R.items()}, init_type,
                       random_state)
    else:
        G = {k: np.array(v, dtype=float) for k, v in G0.items()}
    S = None
    if stopping_system:
        compute_err = True
    err_target = (None, None)
    err_system = (None, None)
    Tp, Tn = _theta_split(Theta)
    R = {key: [m.copy() for m in mats] for key, mats in R.items()}                  # :268

    for it in range(max_iter):
        if it > 1 and stopping and err_target[1] - err_target[0] < stopping[1]:
            break
        if it > 1 and stopping_system and err_system[1] - err_system[0] < stopping_system:
            break
        if it == 0:                                                                  # :287-292
            for key in M:
                for l in range(len(R[key])):
                    if M[key][l] is not None:
                        R[key][l][M[key][l]] = 0.
        S = _solve_backbones(R, G)
        for key in M:                                                                # :319-325
            for l in range(len(M[key])):
                if M[key][l] is None:
                    continue
                ti, tj = key
                approx = np.dot(G[ti, ti], np.dot(S[ti, tj][l], G[tj, tj].T))
                R[key][l][M[key][l]] = approx[M[key][l]]
        num = {key: np.zeros(Gt.shape) for key, Gt in G.items()}
        den = {key: np.zeros(Gt.shape) for key, Gt in G.items()}
        for (ti, tj), mats in R.items():
            for l, mat in enumerate(mats):
                (ni, di), (nj, dj) = _relation_terms(mat, G[ti, ti], G[tj, tj], S[ti, tj][l], False)
                num[ti, ti] += ni
                den[ti, ti] += di
                num[tj, tj] += nj
                den[tj, tj] += dj
        for key, mats in Tp.items():
            for th in mats:
                den[key] += np.dot(th, G[key])
        for key, mats in Tn.items():
            for th in mats:
                num[key] += np.dot(th, G[key])
        _apply_update(G, num, den)

        if stopping:
            (key, l), _eps = stopping
            approx = np.dot(G[key[0], key[0]], np.dot(S[key][l], G[key[1], key[1]].T))
            err_target = (np.linalg.norm(R[key][l] - approx), err_target[0])
        if compute_err:
            total, _ = objective(R, G, S)
            if history is not None:
                history.append(total)
            if stopping_system:
                err_system = (total, err_system[0])
        if callback:
            callback(G, S, it)
    return G, S


# --------------------------------------------------------------------------- transform
def transform(R_ij, Theta_i, target_obj_type, obj_type2rank, G, S, max_iter=10, init_type="random_c",
              stopping=None, stopping_system=None, verbose=0, compute_err=False, callback=None,
              random_state=None, G0=None, history=None):
    """decomposition/_dfmf.py:330-458: only the target factor moves; G (other types) and S frozen.
    Types are matched by identity (``is``) as in _dfmf.py:392,407."""
    rs = random_state if isinstance(random_state, np.random.RandomState) else np.random.RandomState(random_state)
    tgt = target_obj_type
    sizes = [mats[0].shape[0 if tgt == ti else 1] for (ti, tj), mats in R_ij.items()]
    n_new = sizes[0]
    if G0 is None:
        Gx = initialize([tgt], {tgt: n_new}, obj_type2rank, {k: v[0] for k, v in R_ij.items()}, init_type, rs)
        Gi = Gx[tgt, tgt]
    else:
        Gi = np.array(G0, dtype=float)
    if stopping_system:
        compute_err = True
    err_system = (None, None)
    Tp, Tn = [], []
    for mats in Theta_i.values():
        for th in mats:
            p, n = _sign_split(th)
            Tp.append(p)
            Tn.append(n)

    for it in range(max_iter):
        if it > 1 and stopping_system and err_system[1] - err_system[0] < stopping_system:
            break
        num = np.zeros(Gi.shape)
        den = np.zeros(Gi.shape)
        for (ti, tj), mats in R_ij.items():
            for l, mat in enumerate(mats):
                Sl = S[ti, tj][l]
                if ti is tgt:                                                        # :392-405
                    t1p, t1n = _sign_split(np.dot(mat, np.dot(G[tj, tj], Sl.T)))
                    t2p, t2n = _sign_split(np.dot(Sl, np.dot(G[tj, tj].T, np.dot(G[tj, tj], Sl.T))))
                    num += t1p + np.dot(Gi, t2n)
                    den += t1n + np.dot(Gi, t2p)
                if tj is tgt:                                                        # :407-419
                    t4p, t4n = _sign_split(np.dot(mat.T, np.dot(G[ti, ti], Sl)))
                    t5p, t5n = _sign_split(np.dot(Sl.T, np.dot(G[ti, ti].T, np.dot(G[ti, ti], Sl))))
                    num += t4p + np.dot(Gi, t5n)
                    den += t4n + np.dot(Gi, t5p)
        for th in Tp:
            den += np.dot(th, Gi)
        for th in Tn:
            num += np.dot(th, Gi)
        Gi = np.multiply(Gi, np.sqrt(np.divide(num, np.maximum(den, EPS64))))
        if compute_err:
            total = 0.
            for (ti, tj), mats in R_ij.items():
                for l, mat in enumerate(mats):
                    if ti is tgt:
                        approx = np.dot(Gi, np.dot(S[ti, tj][l], G[tj, tj].T))
                    else:
                        approx = np.dot(G[ti, ti], np.dot(S[ti, tj][l], Gi.T))
                    total += np.linalg.norm(mat - approx, "fro")
            if history is not None:
                history.append(total)
            if stopping_system:
                err_system = (total, err_system[0])
        if callback:
            callback(Gi, it)
    return Gi


# --------------------------------------------------------------------------- synthetic workloads
def synthetic_graph(n, n_types=5, rank=64, seed0=1000, storage="float64"):
    """The benchmark graph of SURVEY.md §8(d): types 0..T-1, one relation per pair i<j,
    R_ij = RandomState(seed0 + 10*i + j).rand(n, n) rounded to the storage dtype and handed back
    as float64 (so CPU and GPU see identical numbers).  storage: float64 | float32 | bfloat16."""
    types = list(range(n_types))
    R = {}
    for i in types:
        for j in types:
            if i < j:
                m = np.random.RandomState(seed0 + 10 * i + j).rand(n, n)
                R[i, j] = [round_to_storage(m, storage)]
    ranks = {t: rank for t in types}
    return types, ranks, R


def hashed_uniform(seed, rows, cols, row0=0):
    """numpy twin of the engine's fz_fill_uniform (csrc/fz_kernels.cuh:hashed_uniform): float64 array of
    rows x cols values k / 2^24, element (r, c) a pure function of (seed, (row0 + r) * cols + c)."""
    idx = (np.arange(row0, row0 + rows, dtype=np.uint64)[:, None] * np.uint64(cols) + np.arange(cols, dtype=np.uint64)[None, :])
    with np.errstate(over="ignore"):
        z = np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + idx
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(40)).astype(np.float64) / 16777216.0


def hashed_graph(n, n_types=5, rank=64, seed0=1000, storage="bfloat16"):
    """The benchmark graph with counter-based entries: R_ij = hashed_uniform(seed0 + 10 i + j) for i < j."""
    types = list(range(n_types))
    R = {(i, j): [round_to_storage(hashed_uniform(seed0 + 10 * i + j, n, n), storage)]
         for i in types for j in types if i < j}
    return types, {t: rank for t in types}, R


def round_to_storage(a, storage):
    a = np.asarray(a, dtype=np.float64)
    if storage == "float64":
        return a
    if storage == "float32":
        return a.astype(np.float32).astype(np.float64)
    if storage == "bfloat16":
        return bf16_round(a)
    raise ValueError(storage)


def bf16_round(a):
    """Round-to-nearest-even to bfloat16, returned as float64 (matches __float2bfloat16_rn on fp32)."""
    f = np.asarray(a, dtype=np.float32)
    u = f.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    out = rounded.astype(np.uint32).view(np.float32)
    out = np.where(np.isfinite(f), out, f)
    return out.astype(np.float64)
