"""A non-Python host of the C ABI: tests/cabi_smoke.c (built by neuralquantum.jl_b200/build.py with gcc) drives
create -> set_params -> logpsi_grad -> local_scalar -> center -> force -> sr_setup -> solve -> update with HOST buffers
and checks every step in C (finite differences, flipped configurations, direct sums, residuals)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cabi_smoke")


def test_c_host_is_built_and_links_the_library():
    assert os.path.exists(EXE), "run python neuralquantum.jl_b200/build.py"
    out = subprocess.run(["ldd", EXE], capture_output=True, text=True).stdout
    assert "libnqcuda.so" in out and "not found" not in out.split("libnqcuda.so")[1].splitlines()[0]


@pytest.mark.gpu
def test_c_host_runs_the_iteration_with_host_buffers():
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "CABI_SMOKE_OK" in r.stdout, r.stdout + r.stderr
