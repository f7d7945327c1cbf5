"""world_size-2 gloo test (CPU) of the row-sharded hot loop: the partition arithmetic, the packed fp64
all-reduce, the per-relation reduce-scatter and the factor all-gather of
skfusion.fusion.distributed.run_iterations, driven with a numpy stand-in for the per-rank engine."""
import os
import socket
import warnings

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import cases
import fusion_oracle as oracle
from helpers import rel_fro


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, case_name, n_iters, out_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    for p in (here, os.path.join(here, "golden"), os.path.join(root, "oracle"), os.path.join(root, "scikit-fusion_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from numpy_shard import NumpyShard
    from skfusion.fusion.distributed import Collectives, local_rows, run_iterations
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = cases.fit_cases()[case_name]
    sizes = oracle.count_objects(case["R"])
    G0 = oracle.initialize(case["types"], sizes, case["ranks"], {}, "random", np.random.RandomState(case["seed"]))
    R_local = {}
    for (ti, tj), mats in case["R"].items():
        lo, hi = local_rows(sizes[ti], world, rank)
        R_local[ti, tj] = [m[lo:hi] for m in mats]
    shard = NumpyShard(R_local, case["types"], sizes, case["ranks"], G0, world, rank)
    run_iterations(shard, Collectives(dist), n_iters)
    G, S = shard.result()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **{"G_%s" % t: G[t, t] for t in case["types"]},
             **{"S_%s_%s_%d" % (k[0], k[1], l): s for k, v in S.items() for l, s in enumerate(v)})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case_name,world", [("readme3", 2), ("multi_theta_no_theta", 2), ("readme3", 3)])
def test_sharded_loop_matches_oracle(tmp_path, case_name, world):
    base = case_name.replace("_no_theta", "")
    case = dict(cases.fit_cases()[base])
    n_iters = 8
    port = _free_port()
    mp.spawn(_worker, args=(world, port, base, n_iters, str(tmp_path)), nprocs=world, join=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Go, So = oracle.dfmf(case["R"], {}, case["types"], case["ranks"], max_iter=n_iters, init_type="random",
                             random_state=np.random.RandomState(case["seed"]))
    for rank in range(world):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        for t in case["types"]:
            assert rel_fro(Go[t, t], got["G_%s" % t]) < 1e-9      # every rank ends with the full factors
        for (ti, tj), mats in So.items():
            for l, s in enumerate(mats):
                assert rel_fro(s, got["S_%s_%s_%d" % (ti, tj, l)]) < 1e-8


def test_regrouped_algebra_equals_reference_order_unsharded():
    """world 1: the two-product regrouping (F6) alone, against the oracle's three-product order."""
    from numpy_shard import NumpyShard
    from skfusion.fusion.distributed import run_iterations

    class Solo(object):
        world, rank = 1, 0
    case = cases.fit_cases()["multi_theta"]
    sizes = oracle.count_objects(case["R"])
    G0 = oracle.initialize(case["types"], sizes, case["ranks"], {}, "random", np.random.RandomState(7))
    shard = NumpyShard(case["R"], case["types"], sizes, case["ranks"], G0, 1, 0)
    run_iterations(shard, Solo(), 25)
    G, S = shard.result()
    Go, So = oracle.dfmf(case["R"], {}, case["types"], case["ranks"], max_iter=25, G0=G0)
    for t in case["types"]:
        assert rel_fro(Go[t, t], G[t, t]) < 1e-10
    for key in So:
        for l, s in enumerate(So[key]):
            assert rel_fro(s, S[key][l]) < 1e-9


def test_local_rows_partition_is_exact_and_contiguous():
    from skfusion.fusion.distributed import local_rows
    for n in (1, 7, 8, 9, 100, 1219):
        for world in (1, 2, 3, 8):
            spans = [local_rows(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) == (n + world - 1) // world
