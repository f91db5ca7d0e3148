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


def test_reference_arm_runs_the_reference_itself_on_its_own_case():
    """BASELINE config 1 (--workload sod) is the one case the reference can run: the arm then times the
    reference's own reader + Time::goNextTimeStep + RhoSolver (oracle/_ref/ref_io), kind "reference"."""
    import pytest
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_io")):
        pytest.skip("oracle/_ref/ref_io not built (make -C oracle ref)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "sod",
                          "--steps", "5", "--warmup", "3"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 8
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and "2d-shockwavepipe-2" in d["config"]["workload"]


def test_our_arm_has_no_cpu_path():
    from conftest import have_gpu
    if have_gpu():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "sod", "--steps", "2"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)
