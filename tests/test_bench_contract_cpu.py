"""bench.py contract checks that need no GPU: the reference arm prints one valid JSON line (the unmodified reference
from oracle/_ref - or /root/reference - on a bounded sample, the oracle port when no copy is present), and the GPU arms
refuse to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=600, env=env)


def _reference_line(env=None):
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-utts", "1", "--nsteps-denoiser", "2",
              "--nsteps-durgen", "2", "--utterances", "8"], env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line():
    have_ref = os.path.isdir("/root/reference/flamed") or os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "flamed"))
    d = _reference_line()
    assert d["impl"] == "reference" and d["unit"] == "audio_s/s" and d["higher_is_better"] is True
    assert d["metric"] == "audio_seconds_per_second_at_128_denoiser_steps"
    assert d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "audio_s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert "fp32" in d["config"]["precision"] and "CPU" in d["config"]["noise"]  # the arm describes itself, not ours


def test_reference_arm_falls_back_to_the_port_without_a_reference_copy(tmp_path):
    d = _reference_line(env=dict(os.environ, FLAMED_REFERENCE_ROOT=str(tmp_path)))
    assert d["cpu_baseline"]["kind"] == "port" and d["value"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    for impl in ("ours", "eager"):
        r = _run(["--impl", impl, "--steps", "1", "--warmup", "0"])
        assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
