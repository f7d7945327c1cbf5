"""float32 relations at scale: exact CUDA-core path vs bf16 planes on the tensor cores (storage='bfloat16x3') vs bf16 storage,
plus ranks above 64.  One JSON line per configuration (it/s with the relations resident in HBM, CUDA events around the
iterations; relFro of the factors against the float32 CUDA-core run of the same seeds).  Verdict row J2.
    python scripts/x3_bench.py [n] [iters] [comma-separated workload labels]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-fusion_b200"))


def main():
    import torch
    from skfusion import _capi
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    only = set(sys.argv[3].split(",")) if len(sys.argv) > 3 else None
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    n_types = 3
    pairs = [(i, j) for i in range(n_types) for j in range(n_types) if i < j]
    R32 = {p: torch.rand((n, n), generator=gen, device=dev, dtype=torch.float32) for p in pairs}
    ratings = {p: torch.randint(0, 6, (n, n), generator=gen, device=dev).to(torch.float32) for p in pairs}
    theta = torch.where(torch.rand((n, n), generator=gen, device=dev) < 0.001, -0.005, 0.0).to(torch.float32)
    unknown = (torch.rand((n, n), generator=gen, device=dev) < 0.3).to(torch.uint8)      # dfmc: 30 % of relation (0, 1) unknown
    results = {}

    def run(label, data, storage, rank, with_theta=False, terms="auto", algo=None):
        algo = _capi.FZ_DFMF if algo is None else algo
        eng = _capi.Engine(device=0, compute="float32")
        try:
            eng.set_split_terms(terms)
            tids = [eng.add_type(n, rank) for _ in range(n_types)]
            for (i, j) in pairs:
                mat = data[i, j]
                if storage == "bfloat16":
                    mat = mat.to(torch.bfloat16)
                masked = algo == _capi.FZ_DFMC and (i, j) == pairs[0]
                eng.add_relation(tids[i], tids[j], mat, storage=storage, borrow=not masked, mask=unknown if masked else None)
            if with_theta:
                eng.add_relation(tids[0], tids[0], theta, storage=storage if storage == "bfloat16x3" else None, borrow=True)
            rs = np.random.RandomState(0)
            for t in tids:
                eng.set_factor(t, rs.rand(n, rank).astype(np.float32))
            eng.finalize()
            eng.iterate(algo, 3)
            torch.cuda.synchronize()
            l0 = eng.launches
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.iterate(algo, iters, torch.cuda.current_stream().cuda_stream)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / iters
            G = [eng.get_factor(t) for t in tids]
            entries = float(len(pairs)) * n * n
            line = {"workload": label, "n": n, "rank": rank, "storage": storage or "float32", "constraint": bool(with_theta),
                    "it_per_s": round(1e3 / ms, 2), "ms_per_it": round(ms, 3),
                    "relation_entries_per_s": round(entries / ms * 1e3, 1), "launches_per_it": (eng.launches - l0) / iters,
                    "operand_stats": eng.operand_stats()}
            results[label] = G
            return line
        finally:
            eng.close()

    def rel(a, b):
        return float(np.linalg.norm(a - b) / np.linalg.norm(a))

    lines = []
    for label, data, storage, rank, th, algo in [(c + (None,) if len(c) == 5 else c) for c in [
            ("f32 cuda-core", R32, None, 64, False),
            ("f32 planes", R32, "bfloat16x3", 64, False),
            ("f32 rounded to bf16", R32, "bfloat16", 64, False),
            ("f32 cuda-core + constraint", R32, None, 64, True),
            ("f32 planes + constraint", R32, "bfloat16x3", 64, True),
            ("ratings cuda-core", ratings, None, 64, False),
            ("ratings planes", ratings, "bfloat16x3", 64, False),
            ("rank128 cuda-core", R32, None, 128, False),
            ("rank128 planes", R32, "bfloat16x3", 128, False),
            ("completion cuda-core", R32, None, 64, False, _capi.FZ_DFMC),
            ("completion planes", R32, "bfloat16x3", 64, False, _capi.FZ_DFMC)]]:
        if only is not None and label not in only:
            continue
        t0 = time.time()
        line = run(label, data, storage, rank, th, algo=algo)
        base = {"f32 planes": "f32 cuda-core", "f32 rounded to bf16": "f32 cuda-core", "f32 planes + constraint": "f32 cuda-core + constraint",
                "ratings planes": "ratings cuda-core", "rank128 planes": "rank128 cuda-core", "completion planes": "completion cuda-core"}.get(label)
        if base and base in results:
            line["relFro_G_vs_cuda_core"] = max(rel(a, b) for a, b in zip(results[base], results[label]))
        line["wall_s"] = round(time.time() - t0, 1)
        print(json.dumps(line), flush=True)
        lines.append(line)


if __name__ == "__main__":
    main()
