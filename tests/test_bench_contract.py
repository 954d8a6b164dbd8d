"""CPU: the driver-facing contract of bench.py that can be checked without a GPU — the `--impl reference` arm (the reference's
own modules from oracle/_ref on the host cores, else the oracle port) prints ONE JSON line with the agreed keys, and only on rank 0."""
import json
import os
import subprocess
import sys

from conftest import ROOT

REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def run(extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--tokens", "4", "--nfe", "1"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.strip()]


def test_reference_arm_json_line():
    lines = run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "audio_seconds_per_second" and d["unit"] == "audio-s/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"]
    built = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "jyutvoice", "__init__.refbin")) or os.path.isdir("/root/reference/jyutvoice")
    assert d["cpu_baseline"]["kind"] == ("reference" if built else "port")  # the reference's own modules when oracle/_ref exists
    assert d["config"]["precision"] == "fp32" and d["dtype"] == "f32" and d["p50_ms"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_is_rank0_only():
    assert run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
