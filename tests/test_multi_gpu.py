"""Sharded DFMF on real GPUs (skipped on boxes with fewer than two): one process per GPU under torch.distributed.run with
the collectives inside the library and on the host, and the whole box driven from ONE process through the reference-facing
API -- solver.dfmf(..., n_gpus=2) / Dfmf(n_gpus=2).fuse(graph) (reference entry: decomposition/dfmf.py:55-106)."""
import os
import subprocess
import sys
import warnings

import numpy as np
import pytest

import fusion_oracle as oracle
from helpers import rel_fro

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("peer_rs", ["1", "0"])
def test_sharded_fit_under_torchrun_matches_the_oracle(peer_rs):
    """peer_rs = 1: the B partials are exchanged by the engine's pull kernel over NVLink peer memory (CUDA IPC between the two
    processes); 0: by NCCL's reduce-scatter.  Same parity bar either way."""
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "run_sharded_nccl_check.py")],
                         capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, FZ_PEER_RS=peer_rs, FZ_GATE_LOG="1"))
    assert ("reduce-scatter over NVLink peer memory" in out.stderr) == (peer_rs == "1"), out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("sharded NCCL check")]
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert len(lines) == 8 and all(l.endswith("PASS") for l in lines), "\n".join(lines)


@pytest.mark.parametrize("storage,dtype,terms,n,tol_g,tol_s", [("bfloat16", "float32", 2, 1200, 1e-3, 5e-3),
                                                                 ("bfloat16", "float32", "auto", 1200, 1e-3, 5e-3),
                                                                 (None, "float64", 2, 500, 1e-9, 1e-8)])
def test_one_process_drives_two_gpus_through_the_seam(storage, dtype, terms, n, tol_g, tol_s):
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    from skfusion.fusion import solver
    types, ranks, R = oracle.hashed_graph(n, n_types=3, rank=64 if storage else 24, storage=storage or "float64")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Go, So = oracle.dfmf(R, {}, types, ranks, max_iter=8, init_type="random_vcol", random_state=np.random.RandomState(3))
        G, S = solver.dfmf(R, {}, types, ranks, max_iter=8, init_type="random_vcol", random_state=np.random.RandomState(3),
                           dtype=dtype, storage=storage, split_terms=terms, n_gpus=2)
    assert solver.last_fit_info["n_gpus"] == 2
    assert max(rel_fro(Go[t, t], G[t, t]) for t in types) < tol_g
    assert max(rel_fro(So[k][0], S[k][0]) for k in So) < tol_s


def test_objective_stopping_and_callback_on_two_gpus_follow_the_oracle():
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    from skfusion.fusion import solver
    types, ranks, R = oracle.synthetic_graph(300, n_types=3, rank=12)
    seen, seen_o = [], []
    Go, So = oracle.dfmf(R, {}, types, ranks, max_iter=30, init_type="random", random_state=np.random.RandomState(1),
                         stopping_system=0.05, compute_err=True, callback=lambda G, S, it: seen_o.append(it))
    G, S = solver.dfmf(R, {}, types, ranks, max_iter=30, init_type="random", random_state=np.random.RandomState(1),
                       stopping_system=0.05, compute_err=True, callback=lambda G, S, it: seen.append(it), dtype="float64", n_gpus=2)
    assert seen == seen_o and 2 < len(seen) < 30          # the same early stop, found from the sharded objective
    assert max(rel_fro(Go[t, t], G[t, t]) for t in types) < 1e-8


def test_estimator_keyword_n_gpus():
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    from skfusion import fusion
    rs = np.random.RandomState(0)
    t1, t2 = fusion.ObjectType("a", 8), fusion.ObjectType("b", 6)
    rel = fusion.Relation(rs.rand(90, 70), t1, t2)
    graph = fusion.FusionGraph([rel])
    one = fusion.Dfmf(max_iter=12, init_type="random", random_state=5, dtype="float64").fuse(graph)
    two = fusion.Dfmf(max_iter=12, init_type="random", random_state=5, dtype="float64", n_gpus=2).fuse(graph)
    assert rel_fro(one.factor(t1), two.factor(t1)) < 1e-10 and rel_fro(one.backbone(rel), two.backbone(rel)) < 1e-10


def test_completion_on_two_gpus_equals_the_oracle():
    """Dfmc shards like Dfmf: the imputation of the masked entries is local to the rows a rank holds."""
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    from skfusion.fusion import solver
    rs = np.random.RandomState(2)
    types, ranks = ["u", "m", "g"], {"u": 7, "m": 9, "g": 3}
    R = {("u", "m"): [rs.rand(81, 64)], ("m", "g"): [rs.rand(64, 11)]}
    M = {("u", "m"): [rs.rand(81, 64) < 0.3], ("m", "g"): [None]}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Go, So = oracle.dfmc(R, M, {}, types, ranks, max_iter=12, init_type="random", random_state=np.random.RandomState(4))
        G, S = solver.dfmc(R, M, {}, types, ranks, max_iter=12, init_type="random", random_state=np.random.RandomState(4),
                           dtype="float64", n_gpus=2)
    assert max(rel_fro(Go[t, t], G[t, t]) for t in types) < 1e-8
    assert max(rel_fro(So[k][0], S[k][0]) for k in So) < 1e-7


@pytest.mark.parametrize("init", ["random_c", "random_vcol"])
def test_device_side_initialisation_on_two_gpus_equals_the_host_initialisation(init):
    """The data-driven seeds with their column means computed by the shard group (norms summed over the ranks, means
    all-gathered / all-reduced) against the reference-order host initialisation: same RandomState consumption, same factors."""
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    from skfusion.fusion import solver
    types, ranks, R = oracle.synthetic_graph(333, n_types=3, rank=12)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Gh, Sh = solver.dfmf(R, {}, types, ranks, max_iter=4, init_type=init, random_state=np.random.RandomState(8),
                             dtype="float64", n_gpus=2, device_init=False)
        Gd, Sd = solver.dfmf(R, {}, types, ranks, max_iter=4, init_type=init, random_state=np.random.RandomState(8),
                             dtype="float64", n_gpus=2, device_init=True)
    assert max(rel_fro(Gh[t, t], Gd[t, t]) for t in types) < 1e-9
    assert max(rel_fro(Sh[k][0], Sd[k][0]) for k in Sh) < 1e-7
