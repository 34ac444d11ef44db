"""CPU: the driver-facing contract of bench.py that can be checked without a GPU — the reference arm
prints ONE JSON line with the required keys, and the default arm refuses to run without CUDA."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    for extra, metric in ((["--cpu-sample", "1"], "structures_per_sec_EFS"),
                          (["--workload", "c5", "--cpu-sample", "1"], "train_structures_per_sec_EFSM")):
        res = _run("--impl", "reference", "--steps", "1", "--warmup", "1", *extra)
        assert res.returncode == 0, res.stderr[-500:]
        lines = [ln for ln in res.stdout.strip().splitlines() if ln.startswith("{")]
        assert len(lines) == 1
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and d["metric"] == metric and d["unit"] == "structures/s" and d["value"] > 0
        for key in ("n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                    "cpu_baseline", "e2e"):
            assert key in d, key
        assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
        assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        assert d["vs_baseline"] is None and d["higher_is_better"] is True and "workload" in d["config"]
        assert d["product_so_mapped"] is False  # inputs come from the numpy builder: the product .so is never loaded
        assert {"task", "whole_job", "weights", "l2", "parallelism"} <= set(d["config"])  # same object as the product arm's


def test_default_arm_needs_cuda():
    import torch

    if torch.cuda.is_available():
        return
    res = _run("--steps", "1", "--warmup", "1")
    assert res.returncode != 0 and "no CUDA device" in (res.stderr + res.stdout)


def test_reference_arm_under_torchrun_only_rank0_works():
    """N > 1: the driver launches the reference arm like the product arm (one process per GPU); rank 0 alone runs and prints,
    the other ranks exit 0 without work and without output."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stderr[-500:]
    assert not [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
