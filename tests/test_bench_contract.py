"""bench.py's contract pieces that need no GPU: the reference arm prints the driver's JSON line with the same config keys as the
engine arm (the round-1 verdict's `same_config: false`), the profile-backed `roofline.traffic` only speaks for the launch it was
captured from, and the engine arm refuses to run without a CUDA device (no CPU fallback behind the product path)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    p = _run("--impl", "reference", "--workload", "cartpole", "--steps", "1", "--warmup", "1")
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert set(line["config"]) == {"workload", "num_rollouts", "num_timesteps", "controller"}  # == the engine arm's keys
    assert line["config"]["workload"] == "cartpole_vanilla_N8192_T100"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_engine_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = _run("--steps", "1", "--warmup", "1")
    assert p.returncode != 0 and "no CPU fallback" in (p.stdout + p.stderr)


def test_profile_backed_traffic_only_for_the_same_launch():
    import bench
    t, src = bench._traffic_from_profile("autorally_nn_N32768_T100", {"grid": 147, "block": 672})
    assert src == "profiles/r02_final_autorally_k1_kernels.csv" and 26.2e6 < t < 26.5e6  # algorithmic: 26 214 400 B
    assert bench._traffic_from_profile("autorally_nn_N32768_T100", {"grid": 128, "block": 672}) == (None, None)
    assert bench._traffic_from_profile("cartpole_vanilla_N8192_T100", {"grid": 128, "block": 64}) == (None, None)
    cfg_keys = set(bench._base_config(bench._workload(type("A", (), {"workload": "autorally", "rollouts": None, "timesteps": None})())))
    assert cfg_keys == {"workload", "num_rollouts", "num_timesteps", "controller"}
