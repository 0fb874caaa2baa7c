"""CPU: the bench.py contract that can be checked without a GPU -- the reference arm prints ONE JSON line with the agreed keys,
and our arm refuses to run (no CPU fallback) when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, timeout=600)


@pytest.mark.parametrize("extra,metric,unit", [((), "mel_frames_per_s", "mel-frames/s"),
                                               (("--workload", "indextts_gpt", "--new-tokens", "16"), "mel_tokens_per_s", "tokens/s")])
def test_reference_arm_prints_one_json_line(extra, metric, unit):
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--frames", "64", *extra)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == metric and d["unit"] == unit and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine without a GPU")
def test_our_arm_fails_loudly_without_a_gpu():
    r = _run("--steps", "1", "--no-cpu-baseline")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
