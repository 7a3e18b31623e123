"""bench.py --impl reference: the reference's own CPU implementation of the path (oracle/_ref) timed on the host cores -- runs
without a GPU, so its JSON contract is checked here (the driver computes the GPU/CPU ratio from this line)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def test_reference_arm_prints_one_contract_line(harness):
    if not harness.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines  # stdout carries exactly one JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "photons/s" and d["unit"] == "photons/s" and d["higher_is_better"] is True
    assert d["value"] > 1e3 and d["steps"] == 1 and d["n_gpus"] == 1 and d["dtype"] == "f64"
    assert d["config"]["workload"].startswith("conical")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
