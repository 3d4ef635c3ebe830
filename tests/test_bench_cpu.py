"""Host logic of bench.py that can run without a GPU: the reference arm's JSON line and the CPU baseline object."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "5", "--warmup", "3"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-1000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "env-steps/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("C3")


def test_cpu_baseline_follows_the_selected_workload_and_carries_the_python_reference():
    """cpu_oracle_run sizes itself from the workload selected at call time (it once froze 16 384 lattices at definition time), and the
    baseline object carries the committed timing of the UNMODIFIED Python reference next to the live C-port number."""
    sys.path.insert(0, ROOT)
    import importlib
    import bench
    importlib.reload(bench)
    try:
        bench.select_workload("c1")
        cb, value, _ = bench.cpu_oracle_run(0.2)
        assert " of 1 lattices" in cb["sample"] and cb["kind"] == "port" and value > 0
        bench.select_workload("c2")
        cb, _, _ = bench.cpu_oracle_run(0.2)
        assert " of 4096 lattices" in cb["sample"]
    finally:
        bench.select_workload("c3")
        importlib.reload(bench)
    rp = cb.get("reference_python")
    assert rp is not None and rp["unit"] == "env-steps/s" and 1e3 < rp["value"] < 1e6 and rp["cores"] >= 1
    assert "UNMODIFIED" in rp["what"] and rp["source"].startswith("profiles/reference_cpu_")
    json.dumps(cb)
