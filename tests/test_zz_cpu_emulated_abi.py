"""Host-side glue of the table-driven ops and ``HyGrid.geometry`` on CPU: the bodies of their ``-m gpu`` tests run against
an emulation of the C-ABI entry points they call (tests/emulation/abi_emulation.py, own process because it monkeypatches
torch).  Checks what the Python layer hands to the C ABI, not the kernels."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_python_layer_against_the_emulated_c_abi():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emulation", "abi_emulation.py")], capture_output=True,
                       text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    assert r.stdout.count("\nok ") + r.stdout.startswith("ok ") == 13, r.stdout
    assert "emulated entry points:" in r.stdout
