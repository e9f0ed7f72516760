"""The bench.py JSON-line contract the driver depends on: one JSON object on stdout carrying the required keys,
for the reference arm (CPU, runs here) and for the B200 arm (GPU)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(args, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                       timeout=900, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must hold exactly one line: %r" % r.stdout[:500]
    return json.loads(lines[0])


def test_reference_arm_line():
    # OMP_NUM_THREADS=1 is what torchrun exports: the arm must still use every core
    d = _run(["--impl", "reference", "--config", "1", "--steps", "2", "--warmup", "1"], {"OMP_NUM_THREADS": "1"})
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"].startswith("stars/sec") and d["unit"] == "stars/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"] == {"value": d["value"], "unit": "stars/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_b200_arm_line():
    d = _run(["--config", "1", "--nstar", "64", "--steps", "2", "--warmup", "3"])
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["value"] > 0 and d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["limiter"] == "fp32-issue" and 0 < r["step_frac"] <= r["frac"] <= r["frac_per_pass"]
    af = r["algorithmic_flop_instr"]
    assert af["per_star_mean"] > 0 and af["k_mag_mean"] >= 1 and af["k_flux_mean"] >= 2 and 0 < af["survivor_frac_mean"] <= 1
    assert d["config"] == _run(["--impl", "reference", "--config", "1", "--steps", "1", "--warmup", "0"])["config"]
    assert d["e2e_fit_api"]["value"] > 0 and d["cpu_baseline"]["one_core"]["cores"] == 1
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
