"""Where the reference is present (the build container; never the GPU box) the oracle is cross-checked against it live,
on seeded random shapes that are not among the committed golden vectors.  Runs in its own process because the reference
package is also called ``HyGrid``."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/HyGrid"), reason="the reference only exists in the build container")
def test_oracle_matches_the_live_reference():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "live_check.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "live reference check ok" in r.stdout
