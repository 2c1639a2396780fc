"""bench.py contract checks that need no GPU: the reference arm's JSON line, the native arm failing loudly without a
device (no CPU fallback), and the committed profile artefacts bench.py reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def _run(args, timeout=600, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, capture_output=True, text=True,
                          timeout=timeout, env=e)


@pytest.mark.slow
def test_reference_arm_json_line():
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["impl"] == "reference" and d["unit"] == "patches/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_native_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run(["--gpus", "1", "--steps", "1", "--warmup", "1", "--no-cpu-baseline"], timeout=300)
    assert r.returncode != 0
    assert not any(l.startswith('{"metric') for l in r.stdout.splitlines()), "no number may be printed without the CUDA path"


def test_committed_traffic_profile_matches_its_launch_list():
    """profiles/conv_traffic.json (read by bench.py for roofline.traffic) is what tools/traffic_from_ncu.py derives from
    the committed ncu launch list of the same round."""
    tp = os.path.join(ROOT, "profiles", "conv_traffic.json")
    csvp = os.path.join(ROOT, "profiles", "launches_r02_final.csv")
    if not (os.path.exists(tp) and os.path.exists(csvp)):
        pytest.skip("profile artefacts not committed yet")
    out = os.path.join(ROOT, "gpurun_out", "_traffic_check.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "traffic_from_ncu.py"), csvp, out, "conv_tc_kernel", "conv_tc_fold_kernel",
                        "conv_tc_wgrad_kernel"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a, b = json.load(open(tp)), json.load(open(out))
    for k in ("conv_tc_kernel", "conv_tc_fold_kernel", "conv_tc_wgrad_kernel"):
        assert a[k]["launches"] == b[k]["launches"]
        assert abs(a[k]["dram_bytes_per_launch"] - b[k]["dram_bytes_per_launch"]) <= 1e-6 * b[k]["dram_bytes_per_launch"]
