"""The oracle against fixtures generated from the REAL reference (runs anywhere, no GPU)."""
import warnings

import numpy as np
import pytest

import cases
import fusion_oracle as oracle
from helpers import Recorder, rel_fro

TOL = 1e-11  # float64, identical evaluation order: agreement is at rounding level


@pytest.mark.parametrize("name", list(cases.fit_cases().keys()))
def test_fit_trajectory_matches_reference(golden, name):
    case = cases.fit_cases()[name]
    rec = Recorder(case["snapshots"])
    kw = dict(obj_types=case["types"], obj_type2rank=case["ranks"], max_iter=case["max_iter"], init_type=case["init_type"],
              random_state=np.random.RandomState(case["seed"]), callback=rec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if case["algo"] == "dfmc":
            oracle.dfmc(case["R"], case["M"], case["Theta"], **kw)
        else:
            oracle.dfmf(case["R"], case["Theta"], **kw)
    for it in case["snapshots"]:
        for t in case["types"]:
            assert rel_fro(golden["%s/it%d/G/%s" % (name, it, t)], rec.G[it][t]) < TOL
        for (ti, tj), mats in case["R"].items():
            for l in range(len(mats)):
                assert rel_fro(golden["%s/it%d/S/%s,%s/%d" % (name, it, ti, tj, l)], rec.S[it][ti, tj][l]) < TOL


@pytest.mark.parametrize("name", list(cases.fit_cases().keys()))
def test_initial_factors_match_reference(golden, name):
    case = cases.fit_cases()[name]
    n_obj = oracle.count_objects(case["R"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G0 = oracle.initialize(case["types"], n_obj, case["ranks"], {k: v[0] for k, v in case["R"].items()},
                               case["init_type"], np.random.RandomState(case["seed"]))
    for t in case["types"]:
        np.testing.assert_array_equal(G0[t, t], golden["%s/G0/%s" % (name, t)])


@pytest.mark.parametrize("name", list(cases.transform_cases().keys()))
def test_transform_trajectory_matches_reference(golden, name):
    case = cases.transform_cases()[name]
    fit = cases.fit_cases()[case["fit"]]
    last = max(fit["snapshots"])
    tobj = {t: cases.Tag(t) for t in fit["types"]}
    G = {(tobj[t], tobj[t]): golden["%s/it%d/G/%s" % (case["fit"], last, t)] for t in fit["types"]}
    S = {(tobj[a], tobj[b]): [golden["%s/it%d/S/%s,%s/0" % (case["fit"], last, a, b)]] for (a, b) in fit["R"]}
    R_new = {(tobj[a], tobj[b]): m for (a, b), m in case["R_new"].items()}
    Th = {(tobj[a], tobj[a]): m for (a, _), m in case["Theta"].items()}
    ranks = {tobj[t]: r for t, r in fit["ranks"].items()}
    snaps = {}

    def cb(Gi, it):
        if it in case["snapshots"]:
            snaps[it] = np.array(Gi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        oracle.transform(R_new, Th, tobj[case["target"]], ranks, G, S, max_iter=case["max_iter"], init_type=case["init_type"],
                         random_state=np.random.RandomState(case["seed"]), callback=cb)
    for it in case["snapshots"]:
        assert rel_fro(golden["%s/it%d/G" % (name, it)], snaps[it]) < TOL


def test_bf16_rounding_is_round_to_nearest_even():
    x = np.array([1.0, 1.00390625, 1.005859375, 1.0078125, -2.5, 3.1415926, 1e-40, 65504.0], dtype=np.float64)
    r = oracle.bf16_round(x)
    # values exactly representable stay; ties go to even mantissa
    assert r[0] == 1.0 and r[3] == 1.0078125
    assert r[1] == 1.0            # 1 + 2^-8 is a tie between 1 and 1+2^-7 -> even (1.0)
    assert r[2] == 1.0078125      # 1 + 1.5*2^-8 rounds up
    assert abs(r[5] - 3.140625) < 1e-12
