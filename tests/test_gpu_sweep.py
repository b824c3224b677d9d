"""Randomised parity sweep (scripts/sweep.py): 24 candidates at four spreads (from 0.05 to 2.5 x the
NOMAD bounds) x 10 keyframes, every flavour (iba_global, k = 20, iba_global_stable, plane index, GPR,
no plane, wide radius): counters and index-derived quantities bit-exact, sums to 1e-10, LM blocks and
linearisation against the oracle."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_randomised_sweep_over_all_flavours():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sweep.py"), "10", "24"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "SWEEP PASSED" in r.stdout, r.stdout[-3000:]
    assert r.stdout.count(" OK") == 7
