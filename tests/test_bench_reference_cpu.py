"""The reference arm of bench.py (`--impl reference`: the CPU oracle timed on the host cores) runs without a GPU and
prints the contract's JSON line; under torchrun only rank 0 does the work."""

import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(extra_env=None):
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = ""
    env.update(extra_env or {})
    return subprocess.run(
        [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--layers", "2", "--steps", "1", "--warmup", "0"],
        capture_output=True, text=True, env=env, timeout=600, cwd=ROOT,
    )


def test_reference_arm_prints_the_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"].startswith("forecast series/sec") and d["unit"] == "series/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None
    cpu = d["cpu_baseline"]
    assert cpu["kind"] in ("port", "reference") and cpu["cores"] >= 1 and cpu["sample"]
    assert cpu["value"] == d["value"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_without_work():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29533"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_product_arm_refuses_to_run_without_a_gpu():
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = ""
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "0", "--layers", "2"],
                       capture_output=True, text=True, env=env, timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
