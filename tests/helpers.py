"""Shared helpers of the parity tests."""
import numpy as np


def rel_fro(a, b):
    """||a - b||_F / ||a||_F  (a = oracle / reference)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(a), 1e-300))


def golden_factors(golden, case, it, types):
    return {t: golden["%s/it%d/G/%s" % (case, it, t)] for t in types}


def golden_backbones(golden, case, it, R):
    return {key: [golden["%s/it%d/S/%s,%s/%d" % (case, it, key[0], key[1], l)] for l in range(len(mats))]
            for key, mats in R.items()}


def golden_G0(golden, case, types):
    return {(t, t): golden["%s/G0/%s" % (case, t)] for t in types}


class Recorder(object):
    """callback(G, S, it) that keeps deep copies at the wanted iterations."""

    def __init__(self, wanted):
        self.wanted = set(wanted)
        self.G, self.S = {}, {}

    def __call__(self, G, S, it):
        if it in self.wanted:
            self.G[it] = {k[0]: np.array(v) for k, v in G.items()}
            self.S[it] = {k: [np.array(s) for s in v] for k, v in S.items()}
