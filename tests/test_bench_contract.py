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
    p = _run(["--impl", "reference", "--workload", "toy-1k", "--steps", "2", "--warmup", "1", "--batch", "128"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sampled-subgraphs/sec" and d["unit"] == "subgraphs/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "configs[0]" in d["config"]["workload"]
    assert d["config"]["roots_per_step_per_gpu"] == 128  # the reference arm runs the product arm's roots per step


def test_both_arms_describe_the_same_config():
    """`config` is built by one function from (workload, fanout, roots per step, N): the two arms of one driver run print
    the same dict."""
    sys.path.insert(0, ROOT)
    import bench

    a = bench.parse_args(["--gpus", "2", "--impl", "reference"])
    b = bench.parse_args(["--gpus", "2"])
    assert a.batch == b.batch and a.fanout == b.fanout and a.workload == b.workload
    wl = bench.WORKLOADS[a.workload]
    fan = bench.fanout_of(a, wl)
    assert fan == bench.fanout_of(b, wl) == [15, 10]
    assert bench.workload_config(a.workload, wl, fan, min(a.batch, wl["nodes"] // 2), 2) == \
        bench.workload_config(b.workload, wl, fan, min(b.batch, wl["nodes"] // 2), 2)


def test_reference_arm_runs_on_rank_zero_only():
    p = _run(["--impl", "reference", "--workload", "toy-1k", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        return  # on a GPU box the product arm is what the driver runs
    p = _run(["--workload", "toy-1k", "--steps", "1", "--warmup", "1", "--no-cpu-baseline"])
    assert p.returncode != 0 and p.stdout.strip() == ""
