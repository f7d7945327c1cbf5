"""Multi-GPU parity check of the row-sharded path over real NCCL (one rank per GPU under torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/run_sharded_nccl_check.py

Every rank builds its row blocks of the counter-based synthetic graph, runs skfusion.fusion.distributed.dfmf_sharded --
with the collectives inside the library (NCCL on the engine's stream) and spelled out on the host (torch.distributed
between the engine's phases) -- and rank 0 compares the (replicated) result with the float64 oracle.
tests/test_multi_gpu.py launches this file when the box has at least two GPUs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "scikit-fusion_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

CASES = (  # storage, dtype, n, rank, split_terms, tol_G, tol_S
    ("bfloat16", "float32", 1536, 64, 2, 1e-3, 5e-3),
    ("bfloat16", "float32", 1536, 64, "auto", 1e-3, 5e-3),
    ("bfloat16", "float32", 1100, 64, "centred1", 1e-3, 5e-3),
    (None, "float64", 700, 24, 2, 1e-9, 1e-8),
)


def main():
    import torch
    import torch.distributed as dist
    import fusion_oracle as oracle
    from skfusion.fusion import distributed as fzd
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for storage, dtype, n, rank_k, terms, tol_g, tol_s in CASES:
        types, ranks, R = oracle.hashed_graph(n, n_types=3, rank=rank_k, storage=storage or "float64")
        sizes = {t: n for t in types}
        G0 = oracle.initialize(types, sizes, ranks, {}, "random", np.random.RandomState(0))
        R_local = {}
        for (ti, tj), mats in R.items():
            lo, hi = fzd.local_rows(n, world, rank)
            R_local[ti, tj] = [m[lo:hi] for m in mats]
        Go = So = None
        if rank == 0:
            Go, So = oracle.dfmf(R, {}, types, ranks, max_iter=8, G0=G0)
        for mode in ("library", "host"):
            G, S = fzd.dfmf_sharded(R_local, types, sizes, ranks, G0, 8, dist, device=local, collectives=mode, dtype=dtype,
                                    storage=storage, split_terms=terms)
            if rank == 0:
                eg = max(np.linalg.norm(G[t, t] - Go[t, t]) / np.linalg.norm(Go[t, t]) for t in types)
                es = max(np.linalg.norm(S[k][0] - So[k][0]) / np.linalg.norm(So[k][0]) for k in So)
                good = eg < tol_g and es < tol_s
                ok = ok and good
                print("sharded NCCL check world=%d collectives=%s storage=%s dtype=%s split_terms=%s n=%d: relFro G=%.3g S=%.3g %s" % (
                    world, mode, storage, dtype, terms, n, eg, es, "PASS" if good else "FAIL"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
