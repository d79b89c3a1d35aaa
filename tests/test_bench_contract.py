"""bench.py's reference arm (`--impl reference`) is what the driver runs beside the B200 arm: it needs no GPU, times the unmodified
reference on the host cores, and must print ONE JSON line with the contract's keys (same metric / unit / config as the B200 arm,
`impl`, a `cpu_baseline` describing the run, an `e2e` with no copies).  CPU tier: runs wherever the reference (or the oracle) is built."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def test_reference_arm_prints_the_contract_line(checkers):
    if not (checkers.have_ref() or checkers.have_oracle()):
        pytest.skip("neither oracle/_ref nor the oracle is built")
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--seq", "64"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("tokens/sec BioGPT-base") and d["unit"] == "tokens/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["config"]["workload"].startswith("BioGPT-base q4_0 decode")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and abs(cb["value"] - d["value"]) < 1e-6 * max(1.0, d["value"])
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0 and abs(e["value"] - d["value"]) < 1e-6 * max(1.0, d["value"])
