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


def test_bigvgan_stage_table_arithmetic():
    """The per-stage roofline table of the BigVGAN bench line: FLOPs of the 18 convs, SURVEY 8(d)'s M1 bytes (49 C L elements),
    both fractions and the binding floor, from a made-up per-kernel profile."""
    sys.path.insert(0, ROOT)
    import bench
    import b200tts  # noqa: F401
    from b200tts import config
    cfg = config.BIGVGAN
    pk = {"hbm_gbs": 6500.0, "bf16_tflops_sustained": 1400.0}
    prof = {f"bigvgan.resconv.s{i}": 1.5 for i in range(6)}
    prof.update({f"bigvgan.aa_snake.s{i}": 0.5 for i in range(6)})
    rows = bench.bigvgan_stage_table(cfg, 8, 512, prof, pk)
    assert [r["channels"] for r in rows] == [768, 384, 192, 96, 48, 24]
    assert [r["samples"] for r in rows] == [2048, 8192, 16384, 32768, 65536, 131072]
    assert all(r["ms"] == 2.0 for r in rows)
    # stage 0: 6 convs per kernel size 3 / 7 / 11, C = 768, L = 4 T, batch 8
    flops0 = 8 * 6 * 2.0 * 768 * 768 * (3 + 7 + 11) * 2048
    assert rows[0]["tensor_frac"] == pytest.approx(flops0 / 1400e12 * 1e3 / 2.0, rel=1e-3)
    assert rows[0]["hbm_frac_m1"] == pytest.approx(8 * 49 * 768 * 2048 * 2 / 6500e9 * 1e3 / 2.0, rel=1e-2)
    assert [r["bound"] for r in rows] == ["tensor"] * 4 + ["hbm"] * 2            # C <= 48 is HBM-bound under M1
    work = bench.bigvgan_work(cfg, 8, 512)
    total = sum(r["tensor_frac"] * r["ms"] for r in rows) / 1e3 * 1400e12
    assert total == pytest.approx(work["flops_resconv"], rel=1e-3)
    assert bench.bigvgan_stage_table(cfg, 8, 512, {}, pk) == []                  # nothing profiled, nothing claimed
