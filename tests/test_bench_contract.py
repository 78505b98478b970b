"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys, ranks other
than 0 stay silent, and the B200 arm refuses to run (loudly) without a device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                          timeout=600)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1", "--cpu-grid", "48", "--cpu-lattice", "sample"])
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "f64 lattice updates/s" and d["unit"] == "GLUPS"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None and d["value"] > 0
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sub-lattice" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "1024x1024x1024" in d["config"]["workload"] and "model" not in d["config"]
    # the sample the CPU arm really ran is part of its config
    assert d["config"]["reference_sample"]["lattice"] == [48, 48, 48] and cb["spread"]["repetitions"] == 2
    assert cb["spread"]["min"] <= cb["spread"]["median"] <= cb["spread"]["max"]


def test_reference_arm_weak_scaling_workload():
    """--workload C5 (BASELINE configs[4]): 2048 x 2048 x (256 per GPU), harmonic, weak scaling"""
    r = _run(["--impl", "reference", "--workload", "C5", "--gpus", "8", "--steps", "1", "--warmup", "1", "--cpu-grid", "40",
              "--cpu-lattice", "sample"])
    assert r.returncode == 0, r.stderr
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["scaling"] == "weak" and d["config"]["grid"] == [2048, 2048, 2048] and "Harmonic" in d["config"]["workload"]
    assert d["config"]["workload"].startswith("C5") and d["n_gpus"] == 8


def test_reference_arm_other_ranks_do_nothing():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--cpu-grid", "32", "--cpu-lattice", "sample"], env={"RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_b200_arm_has_no_cpu_fallback():
    r = _run(["--gpus", "1", "--steps", "1", "--warmup", "1", "--grid", "32", "--sweeps", "2", "--no-cpu"])
    assert r.returncode != 0 and "no CUDA device" in r.stderr
