"""bench.py's reference arm runs on the CPU (the reference's dfmf from baseline/_ref, else its oracle port, on bounded
samples) and prints ONE JSON line with the contract's keys; checked here with tiny samples so the CPU suite stays fast.  The GPU arm's line is recorded under
profiles/ (it needs a B200)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--cpu-sizes", "128,256"], capture_output=True, text=True, timeout=600, cwd=ROOT,
                         env=dict(os.environ, OMP_NUM_THREADS="1"))     # what torch.distributed.run exports to its workers
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "it/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("DFMF iterations/sec")
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1000.0) < 1e-6
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and "n=256" in cb["sample"]
    assert cb["kind"] == ("reference" if os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "skfusion")) else "port")
    assert cb["cores"] == min(os.cpu_count(), 64) or cb["cores"] == os.cpu_count()    # the BLAS pool is widened past OMP_NUM_THREADS=1
    assert [p["n"] for p in cb["points"]] == [128, 256] and cb["largest_measured"]["n"] == 256 and "max_rel_residual" in cb["fit"]
    assert d["e2e"] == {"value": d["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["n_per_type"] == 81920 and d["config"]["rank"] == 64 and d["config"]["relations"] == 10


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1", "--cpu-sizes", "128,256"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_recorded_gpu_line_carries_the_roofline_and_baseline_objects():
    d = json.loads(open(os.path.join(ROOT, "profiles", "r01b_bench_n1.json")).read().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert r["traffic"] is not None and 1.0 <= r["traffic"] / r["algorithmic_bytes_per_launch"] < 1.2
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["gpu_launches"] > 0 and "workload" in d["config"]
