"""bench.py on a box without a GPU: the reference arm (the oracle port on the host cores) prints the contract's JSON
line, takes every core it may run on even when the launcher exports OMP_NUM_THREADS=1 (torchrun does), and the
product arm refuses to run without a CUDA device instead of falling back to the CPU."""

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--n", "20000", "--nlist", "256", "--nprobe", "4", "--dim", "64", "--cpu-queries", "8"]


def run_bench(args, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=env,
                          timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = run_bench(["--impl", "reference", "--steps", "2", "--warmup", "1", *SMALL], {"OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["unit"] == "queries/s" and d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32"
    # the config is the product arm's (the driver compares the two lines' configs); the bounded sample is described beside it
    assert "workload" in d["config"] and d["config"]["nq"] == 1024 and d["config"]["nprobe"] == 4 and d["config"]["n"] == 20000
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["unit"] == d["unit"] and "8 of the 1024" in cb["sample"]
    assert 0 < cb["coarse_share"] < 1 and cb["scan_host_GBps"] > 0
    want = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    assert cb["cores"] == want  # not the launcher's OMP_NUM_THREADS=1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = run_bench(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", *SMALL], {"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_product_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = run_bench(["--steps", "1", "--warmup", "0", *SMALL])
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stdout + r.stderr)
