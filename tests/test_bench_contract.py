"""bench.py's reference arm runs without a GPU: exactly one JSON line on stdout with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample-cells", "642", "--levels", "8"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cell_level_updates_per_s" and d["unit"] == "cell-levels/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_traffic_contract_numbers_are_the_surveys():
    from mpas_regent_b200 import traffic
    assert traffic.SURVEY_STEP_UNITS_CANONICAL * 8 == 5144 and traffic.SURVEY_STEP_UNITS_LITERAL * 8 == 4800
