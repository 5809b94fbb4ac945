"""CPU: the reference arm of bench.py (the oracle port timed on host cores) prints the contracted JSON line.
The GPU arm's line is exercised on the GPU box; its key set is checked here against the source."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest


@pytest.mark.parametrize("workload", ["C2pile", "C1"])
def test_reference_arm_prints_the_contract_line(workload):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", workload],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "contact_constraint_iterations_per_second" and d["unit"] == "constraint-iters/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["config"]["workload"].startswith(workload + ":") and d["dtype"] == "f32" and d["vs_baseline"] is None
    assert set(d["config"]) == {"workload", "bodies", "solver_iterations", "dt", "window_first_step"}   # the GPU arm prints the same keys
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_gpu_arm_line_has_every_contract_key():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ('"metric"', '"value"', '"unit"', '"n_gpus"', '"steps"', '"warmup"', '"ms_per_step"', '"higher_is_better"', '"scaling"',
                '"vs_baseline"', '"dtype"', '"data"', '"config"', '"roofline"', '"cpu_baseline"', '"e2e"', '"gpu_launches"', '"clocks"',
                '"bound"', '"achieved"', '"peak"', '"frac"', '"traffic"', '"h2d_bytes_per_step"', '"d2h_bytes_per_step"', '"workload"'):
        assert key in src, key
