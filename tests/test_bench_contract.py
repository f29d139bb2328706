"""bench.py contract, the parts that run without a GPU: the reference arm prints exactly one JSON line with the keys the
driver reads; the native arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          env=dict(os.environ, **(env or {})), timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-chunks", "1", "--gpus", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "poses/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    # the unmodified reference out of the staged archive when oracle/stage_ref.py has run (build() does it), else the port
    staged = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "dh_aug_ref.zip"))
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == ("reference" if staged else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_falls_back_to_the_port_without_the_archive(tmp_path):
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-chunks", "1", "--ref-chunk", "64",
             env={"DHFK_REFERENCE_ROOT": str(tmp_path), "DHFK_REFERENCE_ZIP": str(tmp_path / "none.zip")})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["kind"] == "port" and d["value"] > 0


def test_multi_gpu_line_names_configs4():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-chunks", "1", "--ref-chunk", "64", "--gpus", "8",
             env={"RANK": "0", "WORLD_SIZE": "8"})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["scaling"] == "strong" and d["n_gpus"] == 8
    assert "configs[4]" in d["config"]["workload"] and "16777216" in d["config"]["workload"]
    assert d["config"]["poses_per_gpu"] == 16777216 // 8


def test_reference_arm_other_ranks_exit_quietly():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_native_arm_needs_a_gpu():
    r = _run("--steps", "1", "--warmup", "1", "--no-e2e", "--no-cpu-baseline")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
