"""bench.py contract checks that need no GPU: the reference arm (--impl reference) runs the CPU port of the reference prover
and prints ONE JSON line with the keys the driver reads; the B200 arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, env=e)


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-bn", "8"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "hashes/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("proven MiMC hashes/sec")
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "hashes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None
    # the line states the batch that was actually proven (it used to claim the 2^22 workload while proving 2^16)
    cfg = d["config"]
    assert cfg["bn"] == 8 and cfg["hashes_per_step"] == 256 and "2^8-hash batch" in cfg["workload"]
    assert cfg["named_workload_bn"] == 22 and cfg["same_batch_as_b200_arm"] is False
    assert "2^8-hash batch" in d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["ns_per_fr_mul_single_thread"] > 0 and "CIOS" in d["cpu_baseline"]["multiplier"]


def test_reference_arm_same_batch_when_it_fits():
    """--config 1 (2^10 hashes) fits the budget: the reference arm then proves exactly the named batch"""
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--config", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    assert d["config"]["bn"] == 10 and d["config"]["same_batch_as_b200_arm"] is True and d["config"]["baseline_config"].startswith("config 1")


def test_reference_arm_config2_standalone_sumcheck():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--config", "2", "--bn", "10"])
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    assert d["unit"] == "entries/s" and d["value"] > 0 and "sumcheck" in d["metric"] and d["config"]["bn"] == 10


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-bn", "6", "--gpus", "2"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1", "--warmup", "0", "--bn", "4"])
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
