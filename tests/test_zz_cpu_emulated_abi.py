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


def test_gpu_test_bodies_against_the_emulated_c_abi():
    """Every ``-m gpu`` test that does not need the hardware itself (BASELINE-sized inputs and pinned memory are
    deselected) passes with numpy / the oracle standing in for the kernels: the wrappers, autograd
    Functions, modules and numpy shims hand the C ABI what include/hygrid_b200.h says."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emulation", "run_gpu_tests_on_cpu.py")], capture_output=True,
                       text=True, timeout=1800, cwd=ROOT)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
    assert " passed" in tail and "failed" not in tail and int(tail.split(" passed")[0].split()[-1]) >= 220, tail


def test_smoke_and_bench_step_against_the_emulated_c_abi():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emulation", "smoke_on_cpu.py")], capture_output=True,
                       text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "ok smoke and bench step" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
