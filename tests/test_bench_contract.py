"""CPU: the driver-facing contract of bench.py that can be checked without a GPU - the reference arm prints ONE JSON line
with the agreed keys, and the product arm refuses to run on a box without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, cwd=ROOT, env=e, timeout=600)


def test_reference_arm_prints_the_contract_line():
    p = _run(["--impl", "reference", "--workload", "toy-1k", "--steps", "2", "--warmup", "1", "--cpu-sample-roots", "128"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sampled-subgraphs/sec" and d["unit"] == "subgraphs/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "configs[0]" in d["config"]["workload"]


def test_reference_arm_runs_on_rank_zero_only():
    p = _run(["--impl", "reference", "--workload", "toy-1k", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        return  # on a GPU box the product arm is what the driver runs
    p = _run(["--workload", "toy-1k", "--steps", "1", "--warmup", "1", "--no-cpu-baseline"])
    assert p.returncode != 0 and p.stdout.strip() == ""
