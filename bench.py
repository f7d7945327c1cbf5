#!/usr/bin/env python
"""DFMF iterations/sec on the synthetic 5-type / 10-relation graph (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the CPU arm (oracle port of the reference, rank 0 only)

A step = one multiplicative-update iteration over the whole graph.  Workload: types 0..4, one relation per
pair i<j, n objects per type, rank 64, relations stored in bf16 and generated on the device by the engine's
counter-based generator (identical numbers for every sharding).  The same fixed graph is used at every N
(strong scaling): type rows are sharded over the ranks, one process per GPU, NCCL for the three exchanges.
Default n = 81920: the largest power-of-two-ish size whose 10 relations (134 GB) fit one 180 GB B200;
BASELINE configs[3]'s n=100k (200 GB) needs >= 2 GPUs and can be requested with --size 100000 --gpus >= 2.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "scikit-fusion_b200"))

N_TYPES, RANK, SEED0 = 5, 64, 1000
PAIRS = [(i, j) for i in range(N_TYPES) for j in range(N_TYPES) if i < j]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    # (not "--n": torch.distributed.run's own parser treats it as an ambiguous abbreviation of --nnodes / --nproc-per-node)
    ap.add_argument("--size", dest="n", type=int, default=0, help="objects per type (0 = 81920, shrunk if HBM is short)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--split-terms", default="auto", help="1..3, auto or centred1 (operand form of the factors, include/fz_fusion.h)")
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="seconds of host time the CPU arm may spend on its samples")
    ap.add_argument("--cpu-sizes", default="4096,8192,16384", help="objects per type of the CPU arm's bounded samples")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="synthetic", choices=["synthetic", "readme3", "dicty", "movielens", "transform"],
                    help="synthetic = the contract workload; readme3 / dicty / movielens = the small BASELINE configs C1 / C2 / C3 "
                         "(latency-bound; informational line, N=1 only); transform = config C5 (project 10 000 new "
                         "rows of type 0 against a 100 000-object model, 4 relations; informational line, N=1 only)")
    args = ap.parse_args()
    args.cpu_sizes = tuple(int(x) for x in args.cpu_sizes.split(","))
    if args.split_terms not in ("auto", "centred1"):
        args.split_terms = int(args.split_terms)
    return args


# ------------------------------------------------------------------------------------------------ CPU arm
def _blas_threads():
    """Give the BLAS pool every host core (torch.distributed.run exports OMP_NUM_THREADS=1 to its workers) and report
    the thread count the pool really runs with.  Returns (limiter to keep alive, threads, description)."""
    from threadpoolctl import threadpool_info, threadpool_limits
    want = os.cpu_count() or 1
    limiter = threadpool_limits(limits=want)
    blas = [lib for lib in threadpool_info() if lib.get("user_api") == "blas"]
    used = max([int(lib.get("num_threads", 1)) for lib in blas] or [1])
    desc = "; ".join("%s %s threads=%s" % (lib.get("internal_api"), lib.get("version"), lib.get("num_threads")) for lib in blas)
    return limiter, used, "os.cpu_count()=%d, %s" % (want, desc or "no BLAS pool found")


def _reference_dfmf():
    """The reference's own dfmf() (skfusion/fusion/decomposition/_dfmf.py:127) from baseline/_ref (installed by
    baseline/install_reference.sh, compat edits applied in memory), else the oracle port.  Returns (callable, kind)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref_root, "skfusion", "fusion")) and os.environ.get("FZ_BENCH_CPU_PORT", "0") != "1":
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
            import _load_reference as loader
            loader.REF_ROOT = ref_root
            ref_dfmf = loader.functions()[0]

            def run(R, types, ranks, iters, callback):
                return ref_dfmf(R, {}, types, ranks, max_iter=iters, init_type="random", random_state=np.random.RandomState(0),
                                callback=callback, n_jobs=1)
            return run, "reference"
        except Exception as exc:     # an unusable install must not take the bench down: fall back to the port, say so
            print("[bench] reference under baseline/_ref unusable (%r): timing the oracle port" % (exc,), file=sys.stderr)
    import fusion_oracle as oracle

    def run(R, types, ranks, iters, callback):
        return oracle.dfmf(R, {}, types, ranks, max_iter=iters, init_type="random", random_state=np.random.RandomState(0),
                           callback=callback)
    return run, "port"


def _cpu_graph(n_s):
    """Benchmark-shaped graph for the CPU timing legs: bf16-representable uniform values as float64 (timing does not
    depend on the values; the parity leg uses the engine's own counter-based generator instead)."""
    rng = np.random.default_rng(SEED0)
    types = list(range(N_TYPES))
    R = {}
    for (i, j) in PAIRS:
        m = rng.random((n_s, n_s), dtype=np.float32)
        m.view(np.uint32)[...] &= np.uint32(0xFFFF0000)
        R[i, j] = [m.astype(np.float64)]
    return types, {t: RANK for t in types}, R


def cpu_arm(n_workload, budget_s=150.0, sizes=(4096, 8192, 16384), max_iters=5):
    """The reference's dfmf (float64 numpy, n_jobs=1, all host BLAS threads) on bounded samples of the workload: the same
    5-type / 10-relation graph at several sizes, per-iteration time from the callback hook (iteration 0 dropped, median),
    least-squares fit t(n) = a n^2 + b (three n x n x 64 GEMMs per relation + size-independent k x k work), evaluated at
    the workload size.  The largest directly measured point is reported next to it."""
    import psutil
    t_begin = time.perf_counter()
    limiter, threads, thread_desc = _blas_threads()
    run, kind = _reference_dfmf()
    points = []
    queue = list(sizes)
    while queue:
        n_s = queue.pop(0)
        left = budget_s - (time.perf_counter() - t_begin)
        est = points[-1]["s_per_it"] * (float(n_s) / points[-1]["n"]) ** 2 if points else 0.0
        iters = max_iters
        if points:
            iters = int(min(max_iters, (left - 0.25 * est) // max(est, 1e-9)))     # 0.25 est: generating the inputs
        need_b = 10.0 * n_s * n_s * 8 * 2.5
        if points and (iters < 2 or need_b > psutil.virtual_memory().available):
            # this size does not fit the time or the host memory left: a fit wants at least three points (a residual), so
            # fall back to a size half way to it instead of stopping at two
            mid = ((points[-1]["n"] + n_s) // 2 // 1024) * 1024
            if len(points) < 3 and mid > points[-1]["n"]:
                queue = [mid]
                continue
            break
        types, ranks, R = _cpu_graph(n_s)
        stamps = []
        run(R, types, ranks, iters + 1, lambda G, S, it: stamps.append(time.perf_counter()))
        per_it = np.diff(stamps)
        points.append({"n": n_s, "s_per_it": float(np.median(per_it)), "timed_iterations": int(len(per_it))})
        del R
    ns = np.array([p["n"] for p in points], dtype=np.float64)
    ts = np.array([p["s_per_it"] for p in points])
    if len(points) >= 2:
        A = np.stack([ns ** 2, np.ones_like(ns)], axis=1)
        (a, b), *_ = np.linalg.lstsq(A, ts, rcond=None)
        if a <= 0.0 or b < 0.0:
            a, b = float(ts[-1] / ns[-1] ** 2), 0.0
    else:
        a, b = float(ts[-1] / ns[-1] ** 2), 0.0
    resid = float(np.max(np.abs(a * ns ** 2 + b - ts) / ts))
    t_workload = a * float(n_workload) ** 2 + b
    del limiter
    return {"value": 1.0 / t_workload, "cores": threads, "kind": kind, "points": points,
            "fit": {"a_s_per_n2": float(a), "b_s": float(b), "max_rel_residual": round(resid, 4)},
            "largest_measured": {"n": int(ns[-1]), "it_per_s": round(1.0 / float(ts[-1]), 5)},
            "extrapolated": {"n": int(n_workload), "s_per_it": round(t_workload, 3)},
            "threads": thread_desc,
            "sample": "%s dfmf (float64 numpy, n_jobs=1, %d BLAS threads), same 5-type/10-relation rank-64 graph at %s: %s s/it "
                      "(median, iteration 0 dropped); least-squares t(n) = a n^2 + b (max rel. residual %.1f%%) evaluated at n=%d "
                      "-> %.1f s/it EXTRAPOLATED; largest measured point n=%d: %.4f it/s" % (
                          "reference skfusion._dfmf" if kind == "reference" else "oracle port of", threads,
                          ", ".join("n=%d" % p["n"] for p in points), ", ".join("%.3f" % p["s_per_it"] for p in points),
                          100.0 * resid, n_workload, t_workload, int(ns[-1]), 1.0 / float(ts[-1]))}


def parity_leg(split_terms, iters, n_s=2048):
    """Driver-visible parity: the bench's own GPU code path (same generator, storage, operand form) on the same graph
    shape at a size the float64 oracle finishes in seconds, compared factor by factor and backbone by backbone."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fusion_oracle as oracle
    from skfusion.fusion import solver
    types, ranks, R = oracle.hashed_graph(n_s, N_TYPES, RANK, SEED0, "bfloat16")
    Go, So = oracle.dfmf(R, {}, types, ranks, max_iter=iters, init_type="random", random_state=np.random.RandomState(0))
    G, S = solver.dfmf(R, {}, types, ranks, max_iter=iters, init_type="random", random_state=np.random.RandomState(0),
                       dtype="float32", storage="bfloat16", split_terms=split_terms)
    eg = max(float(np.linalg.norm(G[t, t] - Go[t, t]) / np.linalg.norm(Go[t, t])) for t in types)
    es = max(float(np.linalg.norm(S[k][0] - So[k][0]) / np.linalg.norm(So[k][0])) for k in So)
    return {"n_per_type": n_s, "iterations": iters, "relFro_G_max": eg, "relFro_S_max": es, "tolerance": {"G": 1e-3, "S": 5e-3},
            "ok": bool(eg <= 1e-3 and es <= 5e-3), "oracle_check": _check_sums(Go, So, iters), "engine_check": _check_sums(G, S, iters)}


def _check_sums(G, S, iters):
    """One fp64 fingerprint of a fit: sum of the Frobenius norms of the factors and of the backbones."""
    g = float(sum(np.linalg.norm(np.asarray(v, dtype=np.float64)) for v in G.values()))
    sb = float(sum(np.linalg.norm(np.asarray(m, dtype=np.float64)) for v in S.values() for m in v))
    return {"iterations": int(iters), "sum_fro_G": g, "sum_fro_S": sb}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw,power.limit")

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def sample(self):
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
            self.rows.append([c.strip() for c in out.strip().split(",")])
        except Exception:
            pass

    def run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(0.1)

    def sample_instant_power(self):
        """power.draw is a ~1 s moving average and lags a sub-second timed region; newer drivers also expose the
        instantaneous reading.  Optional: an unknown field just leaves it out."""
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=power.draw.instant",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
            if out.returncode == 0:
                self.instant_w = float(out.stdout.strip().split()[0])
        except Exception:
            pass

    def summary(self):
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k] == "Active" for r in self.rows)]
        def num(col):
            vals = []
            for r in self.rows:
                try:
                    vals.append(float(r[col]))
                except (IndexError, ValueError):
                    pass
            return vals
        power, limit = num(6), num(7)
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "power_w": round(max(power), 1) if power else None,
                "power_instant_w": getattr(self, "instant_w", None),
                "power_limit_w": round(max(limit), 1) if limit else None}


def fused_kernel_active(args):
    return os.environ.get("FZ_NO_FUSED", "0") != "1" and args.split_terms in (2, "auto", "centred1")


# ------------------------------------------------------------------------------------------------ GPU arm
def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from skfusion import _capi
    from skfusion.fusion import distributed as fzd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = args.n
    if n <= 0:
        n = 81920
        free_b, _total = torch.cuda.mem_get_info(dev)
        while 10.0 * n * n * 2 / world + 4e9 > free_b and n > 8192:   # relations + workspaces must fit
            n -= 8192
    lo, hi = fzd.local_rows(n, world, rank)
    stream = torch.cuda.current_stream(dev).cuda_stream

    # ---- the graph, resident in HBM before any timed region
    R_local = {}
    for (i, j) in PAIRS:
        # rows pitched to a multiple of 128 bytes: every 128-byte TMA box row then sits in one L2 line (matters at n = 100 000)
        t = torch.empty((hi - lo, (n + 63) // 64 * 64), dtype=torch.bfloat16, device=dev)[:, :n]
        _capi.fill_uniform(t, SEED0 + 10 * i + j, row0=lo, stream=stream)
        R_local[i, j] = [t]
    rs = np.random.RandomState(0)
    types = list(range(N_TYPES))
    sizes = {t: n for t in types}
    ranks = {t: RANK for t in types}
    G0 = {(t, t): rs.rand(n, RANK).astype(np.float32) for t in types}
    opts = dict(dtype="float32", storage="bfloat16", split_terms=args.split_terms)
    eng, tid, rel_ids = fzd.build_sharded_engine(R_local, sizes, ranks, types, G0, world, rank, local, opts)
    if world > 1:
        fzd.attach_comm(eng, dist)       # NCCL inside the library: the whole sharded iteration is one C call

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    eng.iterate(_capi.FZ_DFMF, args.warmup, stream)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launches
    operand0 = eng.operand_stats()
    eng.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    eng.iterate(_capi.FZ_DFMF, args.steps, stream)
    e1.record()
    if rank == 0:
        sampler.sample_instant_power()   # the host has only enqueued the steps: the GPU is in the middle of the timed region
    barrier()
    sampler.stop_flag = True
    if rank == 0 and not sampler.rows:     # timed region shorter than one nvidia-smi round trip: sample right after it
        sampler.sample()
    elapsed_ms = e0.elapsed_time(e1)
    n_prof, prof_ms, prof_bytes, prof_alg = eng.profile_read()
    eng.profile(False)
    operand = eng.operand_stats()
    timed_single = operand["single"] - operand0["single"]
    timed_two = operand["two_term"] - operand0["two_term"]
    launches = eng.launches - launches0
    if world > 1:
        tt = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tt.item())
        ll = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(ll)
        launches = int(ll.item())
    ms_per_step = elapsed_ms / args.steps
    value = 1000.0 / ms_per_step

    # ---- roofline of the dominant kernel (streamed tensor-core product): HBM-bound.  Algorithmic
    # bytes of the pair of products of one relation = ONE pass over its bf16 matrix: a fused launch is credited
    # with the bytes it streams, a single-product launch with half of them.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    alg_bytes_per_launch = prof_alg / max(1, n_prof)
    traffic, traffic_note = None, None
    kernel_name = "umma_fused_t_kernel" if os.environ.get("FZ_FUSED_VER", "3") == "4" else "umma_fused_kernel"
    if timed_single > 0 and timed_single >= timed_two:
        kernel_name = "umma_fused1_kernel"
    try:   # dram__bytes_read+write per launch from the committed ncu --set full capture, scaled to this launch size
        cap = None
        for name in ("r02_ncu_traffic.json", "r01b_ncu_traffic.json"):
            table = json.load(open(os.path.join(ROOT, "profiles", name))) if os.path.exists(os.path.join(ROOT, "profiles", name)) else {}
            if kernel_name in table:
                cap = table[kernel_name]
                break
        if fused_kernel_active(args) and cap is not None:
            traffic = cap["traffic_over_algorithmic"] * alg_bytes_per_launch
            traffic_note = "ncu capture at n=40960: %.4f x algorithmic bytes (%s), scaled to this launch" % (
                cap["traffic_over_algorithmic"], cap["raw"])
    except Exception:
        pass
    avg_ms = prof_ms / max(1, n_prof)
    achieved = alg_bytes_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    fused = fused_kernel_active(args)
    roofline = {"bound": "hbm", "kernel": "umma_fused_kernel (tcgen05/TMA, A and B from one stream of R)" if fused else
                "umma_skinny_kernel<N,*> (tcgen05/TMA, one product per pass)",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src, "launches_timed": n_prof,
                "avg_launch_ms": round(avg_ms, 4), "algorithmic_bytes_per_launch": alg_bytes_per_launch,
                "streamed_bytes_per_launch": prof_bytes / max(1, n_prof),
                "kernel_share_of_step": round(prof_ms / max(1e-9, e0.elapsed_time(e1)), 4)}
    if fused:
        # What actually limits the kernel: with the 2-term split it EXECUTES 2 x 2 x 128 flop per
        # relation element = 256 flop/B, above the measured ridge (sustained bf16 / HBM), and under random operands the
        # board sits at its power cap.  Reported next to the algorithmic (HBM) roofline, never instead of it.
        tf_peak = float(peaks.get("bf16_tflops_sustained", 0.0)) or None
        # 2 products x 2 flop x 64 columns per relation element and split term, weighted by the kernels that actually ran
        terms_run = (timed_single + 2.0 * timed_two) / max(1, timed_single + timed_two)
        executed = 2.0 * 128.0 * terms_run * (alg_bytes_per_launch / 2.0)      # flop per launch
        tf = executed / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
        roofline["kernel"] = kernel_name + " (tcgen05/TMA, A and B from one stream of R)"
        roofline["tensor_executed"] = {"achieved": round(tf, 1), "peak": tf_peak, "unit": "TFLOP/s",
                                       "frac": round(tf / tf_peak, 4) if tf_peak else None,
                                       "note": "executed tensor work (128 flop per relation byte and split term; %.2f terms on average "
                                               "over the timed iterations); peak = MEASURED_PEAKS.json bf16_tflops_sustained" % terms_run}

    # ---- e2e: the same fit through the C ABI with HOST buffers (pinned), H2D of the relations and D2H of the
    # factors / backbones inside the timed region.
    # ---- fingerprint of the state after exactly warmup + steps iterations: identical for every sharding of the graph
    Gf = {(t, t): eng.get_factor(tid[t]) for t in types}
    Sf = {key: [eng.get_backbone(i) for i in ids] for key, ids in rel_ids.items()}
    check = _check_sums(Gf, Sf, args.warmup + args.steps)
    del Gf, Sf
    e2e = None
    eng.close()                      # frees the engine's buffers and drops its references to the borrowed relations
    if not args.no_e2e:
        e2e = e2e_leg(args, torch, dist, fzd, _capi, R_local, G0, types, sizes, ranks, world, rank, local, dev, opts, n, lo, hi)

    out = None
    if rank == 0:
        cpu, parity = None, None
        if world == 1 and not args.no_cpu:
            parity = parity_leg(args.split_terms, args.warmup + args.steps)
            c = cpu_arm(n, budget_s=min(args.cpu_budget, 120.0), sizes=args.cpu_sizes)
            cpu = {"value": c["value"], "unit": "it/s", "cores": c["cores"], "kind": c["kind"], "sample": c["sample"],
                   "points": c["points"], "fit": c["fit"], "largest_measured": c["largest_measured"], "threads": c["threads"]}
        out = {
            "metric": "DFMF iterations/sec on the synthetic 5-type 10-relation graph", "value": round(value, 4), "unit": "it/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "5 object types x n=%d objects, 10 relations (all pairs i<j) n x n stored bf16, rank 64, "
                                   "init 'random', factor operand form split_terms=%s with fp32 accumulation, fp64 k x k chain; "
                                   "type rows sharded over %d GPU(s)" % (n, args.split_terms, world),
                       "n_per_type": n, "rank": RANK, "relations": len(PAIRS), "split_terms": args.split_terms,
                       "relation_bytes_total": 10 * n * n * 2,
                       "cache": "inputs (%.1f GB per GPU) exceed the 126 MB L2; no flush needed" % (10.0 * n * (hi - lo) * 2 / 1e9)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "check": check, "parity": parity,
            "operand_form": {"split_terms": args.split_terms, "single_term_iterations": operand["single"],
                             "two_term_iterations": operand["two_term"], "timed_single_term_iterations": timed_single,
                             "timed_two_term_iterations": timed_two, "gate_error_estimate": operand["err"],
                             "gate_cond_estimate": operand["cond"]},
            "clocks": sampler.summary(),
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


def e2e_leg(args, torch, dist, fzd, _capi, R_dev, G0, types, sizes, ranks, world, rank, local, dev, opts, n, lo, hi):
    """K iterations end to end through the public API with HOST inputs: relation row blocks in pinned host memory (bf16),
    factors as host float32; outputs read back to host numpy.  One GPU: solver.dfmf (the reference's seam function);
    several GPUs under torch.distributed.run: distributed.dfmf_sharded, its one-process-per-GPU form."""
    import psutil
    from skfusion.fusion import solver
    need = 10.0 * (hi - lo) * n * 2
    avail = psutil.virtual_memory().available
    if need * 1.15 > avail / max(1, world if world > 1 else 1):
        return {"value": None, "unit": "it/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                "note": "host RAM too small to stage the relations (%.0f GB needed, %.0f GB available)" % (need / 1e9, avail / 1e9)}
    # stage: device -> pinned host (outside the timed region), then drop the device copies
    t_stage = time.perf_counter()
    host = {}
    for key, mats in R_dev.items():
        h = torch.empty(mats[0].shape, dtype=torch.bfloat16, pin_memory=True)
        h.copy_(mats[0])
        host[key] = [h]
    torch.cuda.synchronize(dev)
    if rank == 0:
        print("[bench] staged %.1f GB of relations in pinned host memory in %.1f s" % (need / 1e9, time.perf_counter() - t_stage),
              file=sys.stderr, flush=True)
    for key in list(R_dev.keys()):
        R_dev[key] = None
    R_dev.clear()
    del mats
    import gc
    gc.collect()
    torch.cuda.empty_cache()

    class FixedInit(object):
        """RandomState stand-in that hands the solver the bench's initial factors (host float32), in type order."""
        def __init__(self):
            self.order = iter(types)

        def rand(self, rows, cols):
            t = next(self.order)
            return G0[t, t]

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    if world == 1:
        G, S = solver.dfmf(host, {}, types, ranks, max_iter=args.steps, init_type="random", random_state=FixedInit(),
                           device=local, **opts)
    else:
        G, S = fzd.dfmf_sharded(host, types, sizes, ranks, G0, args.steps, dist, device=local, **opts)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    h2d = need + sum(g.nbytes for g in G0.values())
    d2h = sum(g.nbytes for g in G.values()) + sum(m.nbytes for v in S.values() for m in v)
    host.clear()
    h = None
    gc.collect()
    try:       # hand the pinned staging buffers back to the OS: the CPU arm that follows needs the host memory
        torch._C._host_emptyCache()
    except Exception:
        pass
    return {"value": round(args.steps / dt, 4), "unit": "it/s", "seconds_total": round(dt, 3),
            "h2d_bytes_per_step": int(h2d / args.steps), "d2h_bytes_per_step": int(d2h / args.steps),
            "api": "skfusion.fusion.solver.dfmf" if world == 1 else "skfusion.fusion.distributed.dfmf_sharded",
            "check": _check_sums(G, S, args.steps),
            "note": "one fit of %d iterations through the public API from pinned host buffers: upload of the relation row blocks "
                    "(once per fit, PCIe-bound: %.1f GB per GPU) + iterations + download of all factors and backbones, per rank" % (
                        args.steps, need / 1e9)}


def small_workload(args):
    """BASELINE configs C1 (README 3-type graph) and C2 (dicty): tiny, launch-latency-bound problems.  Prints the
    engine's it/s (fp32 engine, whole loop in one C call, CUDA events) next to the oracle's on the host."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import cases
    import fusion_oracle as oracle
    from skfusion import _capi
    case = {"dicty": cases.dicty_case, "movielens": cases.movielens_case}.get(args.workload, lambda: cases.fit_cases()["readme3"])()
    completion = case.get("algo") == "dfmc"         # C3 is a matrix-completion fit (Dfmc): masked entries re-imputed every iteration
    algo = _capi.FZ_DFMC if completion else _capi.FZ_DFMF
    sizes = oracle.count_objects(case["R"])
    import warnings
    warnings.simplefilter("ignore")
    G0 = oracle.initialize(case["types"], sizes, case["ranks"], {k: v[0] for k, v in case["R"].items()}, case["init_type"],
                           np.random.RandomState(0))
    eng = _capi.Engine(0, "float32")
    tid = {t: eng.add_type(sizes[t], case["ranks"][t]) for t in case["types"]}
    for blocks in (case["R"], case["Theta"]):
        for (a, b), mats in blocks.items():
            for l, m in enumerate(mats):
                mask = case["M"][a, b][l] if (completion and blocks is case["R"]) else None
                eng.add_relation(tid[a], tid[b], m, mask=mask)
    for t in case["types"]:
        eng.set_factor(tid[t], G0[t, t])
    eng.finalize()
    st = torch.cuda.current_stream().cuda_stream
    eng.iterate(algo, max(3, args.warmup), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.launches
    e0.record()
    eng.iterate(algo, args.steps, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = eng.launches - l0
    eng.close()
    stamps = []
    if completion:
        oracle.dfmc(case["R"], case["M"], case["Theta"], case["types"], case["ranks"], max_iter=min(args.steps, 20) + 2, G0=G0,
                    callback=lambda G, S, it: stamps.append(time.perf_counter()))
    else:
        oracle.dfmf(case["R"], case["Theta"], case["types"], case["ranks"], max_iter=args.steps + 2, G0=G0,
                    callback=lambda G, S, it: stamps.append(time.perf_counter()))
    cpu_it = 1.0 / float(np.median(np.diff(stamps)))
    print(json.dumps({"metric": "%s iterations/sec, BASELINE config %s" % ("DFMC" if completion else "DFMF", args.workload), "value": round(1000.0 / ms, 1),
                      "unit": "it/s", "n_gpus": 1, "steps": args.steps, "ms_per_step": round(ms, 4), "dtype": "f32",
                      "gpu_launches": launches, "launches_per_step": launches / float(args.steps),
                      "cpu_baseline": {"value": round(cpu_it, 1), "unit": "it/s", "cores": os.cpu_count(), "kind": "port",
                                       "sample": "oracle %s on the same inputs, all host BLAS threads" % ("dfmc" if completion else "dfmf")},
                      "note": "latency-bound: no roofline claim (SURVEY.md 8d)"}))


def transform_workload(args):
    """BASELINE config C5: project n_new = 10 000 new objects of type 0 into a fitted space of four
    other types with n = 100 000 objects each (R_new,0j: 10k x 100k bf16, 8 GB in total), rank 64, init 'random',
    100 iterations.  The model (G_j, S_0j) is synthetic: timing does not depend on its values.  The engine computes the
    loop-invariant R_new (G_j S^T) once (one tensor-core pass over R_new) and then iterates on n_new x k operands;
    the reference recomputes that product every iteration (_dfmf.py:392-405).  CPU leg: the oracle's transform on a
    bounded sample (1 000 x 10 000 per relation), scaled by the n_new * n product."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fusion_oracle as oracle
    from skfusion import _capi
    n_new, n, iters = (10000, args.n if args.n > 0 else 100000, 100)
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream(dev).cuda_stream
    rs = np.random.RandomState(0)
    eng = _capi.Engine(0, "float32")
    eng.set_split_terms(args.split_terms)
    tgt = eng.add_type(n_new, RANK)
    others = [eng.add_type(n, RANK) for _ in range(4)]
    rels, keep = [], []
    for j, tj in enumerate(others):
        t = torch.empty((n_new, n), dtype=torch.bfloat16, device=dev)
        _capi.fill_uniform(t, SEED0 + 1 + j, stream=st)
        keep.append(t)
        rels.append(eng.add_relation(tgt, tj, t, storage="bfloat16", borrow=True))
    G_new0 = rs.rand(n_new, RANK).astype(np.float32)
    eng.set_factor(tgt, G_new0)
    for tj in others:
        eng.set_factor(tj, rs.rand(n, RANK).astype(np.float32))
    eng.finalize()
    for r in rels:
        eng.set_backbone(r, rs.rand(RANK, RANK) / RANK)

    def one_call():
        eng.set_factor(tgt, G_new0)
        torch.cuda.synchronize(dev)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        eng.transform_prepare(tgt, st)
        e1.record()
        eng.transform_iterate(iters, st)
        e2.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1), e1.elapsed_time(e2)

    for _ in range(max(3, args.warmup)):
        one_call()
    l0 = eng.launches
    eng.profile(True)
    runs = [one_call() for _ in range(5)]
    # transform needs ONE product per relation, so the algorithmic bytes of a launch are the bytes it streams
    n_prof, prof_ms, prof_alg, _ = eng.profile_read()
    eng.profile(False)
    launches = (eng.launches - l0) // 5
    prep_ms = float(np.median([r[0] for r in runs]))
    iter_ms = float(np.median([r[1] for r in runs]))
    total_ms = prep_ms + iter_ms
    G_out = eng.get_factor(tgt)
    assert np.isfinite(G_out).all()
    eng.close()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = prof_alg / max(1e-9, prof_ms * 1e-3) / 1e9

    # CPU leg: oracle transform, float64, reference evaluation order (recomputes R_new (G_j S^T) per iteration)
    cn_new, cn, cit = 1000, 10000, 5
    tags = [oracle_tag(i) for i in range(5)]
    Rc = {(tags[0], tags[1 + j]): [oracle.bf16_round(oracle.hashed_uniform(SEED0 + 1 + j, cn_new, cn))] for j in range(4)}
    Gc = {(tags[1 + j], tags[1 + j]): rs.rand(cn, RANK) for j in range(4)}
    Sc = {(tags[0], tags[1 + j]): [rs.rand(RANK, RANK) / RANK] for j in range(4)}
    rk = {t: RANK for t in tags}
    stamps = []
    oracle.transform(Rc, {}, tags[0], rk, Gc, Sc, max_iter=cit + 1, init_type="random", random_state=np.random.RandomState(1),
                     callback=lambda g, it: stamps.append(time.perf_counter()))
    cpu_it_s = float(np.median(np.diff(stamps)))
    cpu_it_s_full = cpu_it_s * (float(n_new) * n) / (float(cn_new) * cn)
    print(json.dumps({
        "metric": "transform rows*iterations/sec, BASELINE config C5", "value": round(n_new * iters / (total_ms * 1e-3), 1),
        "unit": "rows*it/s", "n_gpus": 1, "iterations": iters, "wall_ms": round(total_ms, 3),
        "prepare_ms": round(prep_ms, 3), "iterate_ms": round(iter_ms, 3), "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "transform: %d new rows of type 0 x 4 relations %d x %d bf16 (%.1f GB), rank 64, split_terms=%s: 2-term bf16 "
                               "split of the frozen operand G_j S^T, %d iterations" % (n_new, n_new, n, 4.0 * n_new * n * 2 / 1e9,
                                                                                     args.split_terms, iters)},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "umma_skinny_kernel<N,false> (R_new (G_j S^T), once per call)",
                     "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "launches_timed": n_prof},
        "cpu_baseline": {"value": round(n_new / cpu_it_s_full, 1), "unit": "rows*it/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": "oracle transform (float64 numpy, recomputes the relation products every iteration) on "
                                   "%d x %d per relation: %.3f s/it, scaled by n_new*n to %.1f s/it" % (cn_new, cn, cpu_it_s, cpu_it_s_full)},
    }))


class oracle_tag(object):
    """Object-type stand-in with identity semantics (transform matches types with ``is``)."""

    def __init__(self, i):
        self.i = i

    def __repr__(self):
        return "type%d" % self.i


def main():
    args = parse()
    if args.workload == "transform":
        transform_workload(args)
        return
    if args.workload != "synthetic":
        small_workload(args)
        return
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        if rank != 0:
            return
        n = args.n if args.n > 0 else 81920
        c = cpu_arm(n, budget_s=args.cpu_budget, sizes=args.cpu_sizes, max_iters=max(2, min(5, args.steps)))
        print(json.dumps({
            "impl": "reference", "metric": "DFMF iterations/sec on the synthetic 5-type 10-relation graph",
            "value": c["value"], "unit": "it/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 / c["value"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "5 object types x n=%d objects, 10 relations (all pairs i<j) n x n stored bf16, rank 64, "
                                   "init 'random' -- the same graph as the GPU arm; CPU arm: the reference's float64 "
                                   "numpy path on all host cores, measured on bounded samples and extrapolated (see cpu_baseline.sample)" % n,
                       "n_per_type": n, "rank": RANK, "relations": len(PAIRS), "relation_bytes_total": 10 * n * n * 2},
            "cpu_baseline": {"value": c["value"], "unit": "it/s", "cores": c["cores"], "kind": c["kind"], "sample": c["sample"],
                             "points": c["points"], "fit": c["fit"], "largest_measured": c["largest_measured"],
                             "extrapolated": c["extrapolated"], "threads": c["threads"]},
            "e2e": {"value": c["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return
    gpu_arm(args)


if __name__ == "__main__":
    main()
