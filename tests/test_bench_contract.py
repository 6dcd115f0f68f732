"""bench.py contract checks that need no GPU: the reference arm (the CPU restatement of the reference path timed on the host
cores) prints ONE JSON line with the agreed keys, and the GPU arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=600):
    env = dict(os.environ, OMP_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, env=env)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    res = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "msda_fwd_bwd_clips_per_s" and d["unit"] == "clips/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    # a timed step is a bounded sample of the clip (1 of 6 layers + mask); the clip time it implies is reported beside it
    assert d["value"] > 0 and abs(d["value"] - 1e3 / d["clip_ms"]) < 1e-6 * d["value"]
    assert abs(d["ms_per_step"] - (d["layer_ms"] + d["mask_ms"])) < 1e-6 * d["ms_per_step"] and d["ms_per_step"] < d["clip_ms"]
    assert abs(d["clip_ms"] - (6 * d["layer_ms"] + d["mask_ms"])) < 1e-6 * d["clip_ms"]
    assert d["config"]["workload"].startswith("R50_ovis_360") and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["dtype"] == "f32"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_device():
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    res = _run(["--steps", "1", "--warmup", "0"], timeout=300)
    assert res.returncode != 0 and "no CPU fallback" in (res.stderr + res.stdout)
