"""Race hunting on the full-size batch: the same 256 frames through bc_pipeline repeatedly (graph replay and
plain launches alternating) must give bit-identical labels and grids every time.  A missed barrier or an
overtaken ring slot in the warp-specialised kernels shows up as a run-to-run difference (or a trap)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_pipeline_is_deterministic_at_full_batch():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_determinism.py"), "60"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "0 differing" in r.stdout
