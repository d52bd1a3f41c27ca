"""tools/sanitize_workload.py touches every kernel family once; here it runs plainly and must exit 0 (under compute-sanitizer it
is the memcheck / racecheck workload -- the tool is closed on this GPU pool, profiles/r2l_sanitizer_unavailable.log)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_every_kernel_family_runs():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_workload.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "sanitize workload done" in r.stdout
