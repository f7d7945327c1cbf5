"""Generate tests/golden/*.npz from the REAL reference (container only).

    python tests/golden/make_golden.py

Runs the unmodified reference algorithm (imported in memory by _load_reference.py, four
compat edits, see there) on small seeded problems and stores inputs, initial factors and the
(G, S) trajectory at fixed iterations.  The fixtures travel to the GPU box; the reference does
not.  Cases are defined in cases.py so the tests rebuild identical inputs.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _load_reference as ref  # noqa: E402
import cases  # noqa: E402


def snapshot_recorder(store, prefix, wanted):
    def cb(G, S, it):
        if it in wanted:
            for (t, _), g in G.items():
                store["%s/it%d/G/%s" % (prefix, it, t)] = np.array(g)
            for (ti, tj), mats in S.items():
                for l, s in enumerate(mats):
                    store["%s/it%d/S/%s,%s/%d" % (prefix, it, ti, tj, l)] = np.array(s)
    return cb


def main():
    warnings.simplefilter("ignore")
    dfmf, dfmc, transform, initialize = ref.functions()
    store = {}
    for name, case in cases.fit_cases().items():
        fn = dfmc if case["algo"] == "dfmc" else dfmf
        kwargs = dict(obj_types=case["types"], obj_type2rank=case["ranks"], max_iter=case["max_iter"],
                      init_type=case["init_type"], random_state=np.random.RandomState(case["seed"]),
                      callback=snapshot_recorder(store, name, case["snapshots"]))
        if case["algo"] == "dfmc":
            G, S = fn(case["R"], case["M"], case["Theta"], **kwargs)
        else:
            G, S = fn(case["R"], case["Theta"], **kwargs)
        # the initial factors, so engines can be fed identical G0 without re-deriving the RNG stream
        n_obj = {}
        for (ti, tj), mats in case["R"].items():
            n_obj.setdefault(ti, mats[0].shape[0])
            n_obj.setdefault(tj, mats[0].shape[1])
        G0 = initialize(case["types"], n_obj, case["ranks"], {k: v[0] for k, v in case["R"].items()},
                        case["init_type"], np.random.RandomState(case["seed"]))
        for (t, _), g in G0.items():
            store["%s/G0/%s" % (name, t)] = np.array(g)
        print("fit case %-16s done" % name)

    for name, case in cases.transform_cases().items():
        fit = cases.fit_cases()[case["fit"]]
        last = max(fit["snapshots"])
        tobj = {t: cases.Tag(t) for t in fit["types"]}       # transform() matches types by identity
        G = {(tobj[t], tobj[t]): store["%s/it%d/G/%s" % (case["fit"], last, t)] for t in fit["types"]}
        S = {(tobj[ti], tobj[tj]): [store["%s/it%d/S/%s,%s/0" % (case["fit"], last, ti, tj)]]
             for (ti, tj) in fit["R"]}
        R_new = {(tobj[ti], tobj[tj]): mats for (ti, tj), mats in case["R_new"].items()}
        Th = {(tobj[t], tobj[t]): mats for (t, _), mats in case["Theta"].items()}
        ranks = {tobj[t]: r for t, r in fit["ranks"].items()}
        snaps = {}

        def cb(Gi, it, snaps=snaps, wanted=case["snapshots"]):
            if it in wanted:
                snaps[it] = np.array(Gi)
        transform(R_new, Th, tobj[case["target"]], ranks, G, S, max_iter=case["max_iter"],
                  init_type=case["init_type"], random_state=np.random.RandomState(case["seed"]), callback=cb)
        for it, g in snaps.items():
            store["%s/it%d/G" % (name, it)] = g
        print("transform case %-10s done" % name)

    out = os.path.join(HERE, "reference_trajectories.npz")
    np.savez_compressed(out, **store)
    print("wrote %s (%d arrays, %.1f KB)" % (out, len(store), os.path.getsize(out) / 1024.))


if __name__ == "__main__":
    main()
