"""Host logic of bench.py that can run without a GPU: the reference arm's JSON line and the experiments leg's
failure handling (every child fails here -- no CUDA device -- and must be reported, never raised)."""
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


def test_experiments_leg_reports_failures_instead_of_raising():
    sys.path.insert(0, ROOT)
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("this test is about the no-GPU failure path")
    import bench
    res = bench.run_experiments()
    assert set(res) >= {"default", "stream_obs", "host_expand"}
    for name, r in res.items():
        assert "error" in r or "skipped" in r, (name, r)
    json.dumps(res)
