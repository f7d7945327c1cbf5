"""A numpy stand-in for one rank's engine in sharded mode (TEST CODE).

Implements the same five-method interface as skfusion.fusion.distributed.CudaShard with the regrouped
algebra of the CUDA engine: A = R[p] G_j, B partial = R[p]^T G_i[p], M = G_i[p]^T A, fp64.
Used by the world_size-2 gloo test to validate the host-side sharding / collective logic on a CPU box,
and (world 1) to show that the regrouping itself equals the reference's evaluation order."""
import numpy as np
import scipy.linalg as spla
import torch

EPS64 = np.finfo(float).eps


def _split(x):
    t = x > 0
    return t * x, (t - 1) * x


class NumpyShard(object):
    def __init__(self, R_local, obj_types, sizes, ranks, G0, world, rank):
        self.types = list(obj_types)
        self.k = {t: int(ranks[t]) for t in self.types}
        self.n = dict(sizes)
        self.world, self.rank = world, rank
        self.m = {t: (self.n[t] + world - 1) // world for t in self.types}
        self.npad = {t: self.m[t] * world for t in self.types}
        self.row0 = {t: rank * self.m[t] for t in self.types}
        self.rows = {t: max(0, min(self.n[t], self.row0[t] + self.m[t]) - self.row0[t]) for t in self.types}
        self.G = {}
        for t in self.types:
            g = np.zeros((self.npad[t], self.k[t]))
            g[:self.n[t]] = G0[t, t]
            self.G[t] = torch.from_numpy(g)
        self.rels = [(key, l, np.asarray(mat, dtype=np.float64)) for key, mats in R_local.items() for l, mat in enumerate(mats)]
        count = sum(self.k[t] ** 2 for t in self.types) + sum(self.k[a] * self.k[b] for (a, b), _, _ in self.rels)
        self._small = torch.zeros(count, dtype=torch.float64)
        self.A = [None] * len(self.rels)
        self.Bfull = [torch.zeros(self.npad[key[1]], self.k[key[0]], dtype=torch.float64) for key, _, _ in self.rels]
        self.Bloc = [torch.zeros(self.m[key[1]], self.k[key[0]], dtype=torch.float64) for key, _, _ in self.rels]
        self.S = {}

    def _loc(self, t):
        return self.G[t].numpy()[self.row0[t]:self.row0[t] + self.rows[t]]

    def products(self):
        parts = []
        for t in self.types:
            gl = self._loc(t)
            parts.append((gl.T @ gl).ravel())
        for idx, ((ti, tj), l, mat) in enumerate(self.rels):
            Gj = self.G[tj].numpy()[:self.n[tj]]
            self.A[idx] = mat @ Gj
            self.Bfull[idx].zero_()
            self.Bfull[idx].numpy()[:self.n[tj]] = mat.T @ self._loc(ti)
            parts.append((self._loc(ti).T @ self.A[idx]).ravel())
        self._small.copy_(torch.from_numpy(np.concatenate(parts)))

    def small(self):
        return self._small

    def bpartials(self):
        return [(full.view(-1), loc.view(-1)) for full, loc in zip(self.Bfull, self.Bloc)]

    def update(self):
        buf = self._small.numpy()
        off = 0
        gram, P = {}, {}
        for t in self.types:
            k = self.k[t]
            gram[t] = np.nan_to_num(buf[off:off + k * k].reshape(k, k))
            P[t] = spla.pinv(gram[t])
            off += k * k
        num = {t: np.zeros((self.rows[t], self.k[t])) for t in self.types}
        den = {t: np.zeros((self.rows[t], self.k[t])) for t in self.types}
        for idx, ((ti, tj), l, mat) in enumerate(self.rels):
            ki, kj = self.k[ti], self.k[tj]
            M = np.nan_to_num(buf[off:off + ki * kj].reshape(ki, kj))
            off += ki * kj
            S = np.nan_to_num(P[ti] @ (M @ P[tj]))
            self.S.setdefault((ti, tj), {})[l] = S
            B = (self.Bloc[idx] if self.world > 1 else self.Bfull[idx]).numpy()[:self.rows[tj]] if self.world > 1 \
                else self.Bfull[idx].numpy()[:self.n[tj]]
            t1p, t1n = _split(np.nan_to_num(self.A[idx] @ S.T))
            t2p, t2n = _split(np.nan_to_num(S @ gram[tj] @ S.T))
            t4p, t4n = _split(np.nan_to_num(B @ S))
            t5p, t5n = _split(np.nan_to_num(S.T @ gram[ti] @ S))
            gi, gj = self._loc(ti), self._loc(tj)
            num[ti] += t1p + gi @ t2n
            den[ti] += t1n + gi @ t2p
            num[tj] += t4p + gj @ t5n
            den[tj] += t4n + gj @ t5p
        for t in self.types:
            new = self._loc(t) * np.sqrt(num[t] / np.maximum(den[t], EPS64))
            self.G[t].numpy()[self.row0[t]:self.row0[t] + self.rows[t]] = new

    def factors(self):
        out = []
        for t in self.types:
            whole = self.G[t].view(-1)
            c = self.m[t] * self.k[t]
            out.append((whole, whole[self.rank * c:(self.rank + 1) * c]))
        return out

    def result(self):
        G = {(t, t): self.G[t].numpy()[:self.n[t]].copy() for t in self.types}
        S = {key: [d[l] for l in sorted(d)] for key, d in self.S.items()}
        return G, S
