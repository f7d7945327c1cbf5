"""Row-sharded engine handles on ONE GPU: two handles (rank 0 / rank 1 of a world of 2) live in the same process
and the three exchanges are emulated with plain tensor arithmetic on the buffers fz_comm_* exposes.  This checks
the sharded code paths of the CUDA engine (local row blocks, B partials, packed k x k buffer, factor chunks)
against the oracle without needing two GPUs; the NCCL plumbing itself is exercised by bench.py --gpus N."""
import warnings

import numpy as np
import pytest

import fusion_oracle as oracle
from helpers import rel_fro

pytestmark = pytest.mark.gpu


def _run(world, storage, dtype, n, iters, terms=2, graph=None):
    import torch
    from skfusion import _capi
    from skfusion.fusion import distributed as fzd
    if graph is not None:
        types, ranks, R = graph
        sizes = oracle.count_objects(R)
    else:
        types, ranks, R = oracle.hashed_graph(n, n_types=3, rank=64 if storage == "bfloat16" else 24, storage=storage or "float64")
        sizes = {t: n for t in types}
    G0 = oracle.initialize(types, sizes, ranks, {}, "random", np.random.RandomState(0))
    shards = []
    for rank in range(world):
        R_local = {}
        for (ti, tj), mats in R.items():
            lo, hi = fzd.local_rows(sizes[ti], world, rank)
            R_local[ti, tj] = [m[lo:hi] for m in mats]
        eng, tid, rel_ids = fzd.build_sharded_engine(R_local, sizes, ranks, types, G0, world, rank, 0,
                                                     dict(dtype=dtype, storage=storage, split_terms=terms))
        sh = fzd.CudaShard(eng, 0, sum(len(v) for v in rel_ids.values()), len(types))
        sh.world, sh.rank = world, rank
        shards.append((eng, tid, rel_ids, sh))
    for _ in range(iters):
        for _, _, _, sh in shards:
            sh.products()
        torch.cuda.synchronize()
        total = sum(sh.small().clone() for _, _, _, sh in shards)           # all-reduce
        for _, _, _, sh in shards:
            sh.small().copy_(total)
        n_rel = shards[0][3].n_relations
        parts = [sh.bpartials() for _, _, _, sh in shards]
        for r in range(n_rel):                                                # reduce-scatter
            summed = sum(parts[p][r][0].clone() for p in range(world))
            cnt = parts[0][r][1].numel()
            for p in range(world):
                parts[p][r][1].copy_(summed[p * cnt:(p + 1) * cnt])
        for _, _, _, sh in shards:
            sh.update()
        torch.cuda.synchronize()
        facs = [sh.factors() for _, _, _, sh in shards]                       # all-gather
        for t in range(len(types)):
            cnt = facs[0][t][1].numel()
            chunks = [facs[p][t][1].clone() for p in range(world)]
            for p in range(world):
                for q in range(world):
                    facs[p][t][0][q * cnt:(q + 1) * cnt].copy_(chunks[q])
        torch.cuda.synchronize()
    eng, tid, rel_ids, _ = shards[-1]
    G = {(t, t): eng.get_factor(tid[t]) for t in types}
    S = {key: [eng.get_backbone(i) for i in ids] for key, ids in rel_ids.items()}
    for e, _, _, _ in shards:
        e.close()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Go, So = oracle.dfmf(R, {}, types, ranks, max_iter=iters, G0=G0)
    return types, G, S, Go, So


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_fp64_engine_matches_oracle(world):
    types, G, S, Go, So = _run(world, None, "float64", 333, 6)
    for t in types:
        assert rel_fro(Go[t, t], G[t, t]) < 1e-9
    for key in So:
        assert rel_fro(So[key][0], S[key][0]) < 1e-8


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_tensor_core_engine_matches_oracle(world):
    types, G, S, Go, So = _run(world, "bfloat16", "float32", 1024, 8)
    for t in types:
        assert rel_fro(Go[t, t], G[t, t]) < 1e-3
    for key in So:
        assert rel_fro(So[key][0], S[key][0]) < 5e-3


def test_counter_based_generator_matches_numpy_twin():
    import torch
    from skfusion import _capi
    t = torch.empty((37, 52), dtype=torch.float32, device="cuda")
    _capi.fill_uniform(t, 1023, row0=11)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(t.cpu().numpy().astype(np.float64), oracle.hashed_uniform(1023, 37, 52, row0=11)[:, :])
    b = torch.empty((64, 40), dtype=torch.bfloat16, device="cuda")
    _capi.fill_uniform(b, 7)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(b.float().cpu().numpy().astype(np.float64), oracle.bf16_round(oracle.hashed_uniform(7, 64, 40)))
