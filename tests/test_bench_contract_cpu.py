"""CPU suite, part 6: bench.py's reference arm prints the contract's JSON line."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-n", "12", "--size", "203"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cell_updates_per_sec" and d["unit"] == "cell-updates/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == dict(value=d["value"], unit=d["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert "box203" in d["config"]["workload"]
