"""The C++ host mirror of the reference's Scheduler / SchedulerProfile / Filter / Scorer / Picker interface
(gateway-api-inference-extension_b200/host/epp_scheduler.hpp): its test program restates the reference's own
scheduler tests (TestSchedule, filter chain, integration routing, prefix completion) on the GPU engine."""
import subprocess

import pytest

import _pkg


def test_host_mirror_compiles_and_links():
    exe = _pkg.load_build().build_host_test()
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libeppscore.so" in out and "not found" not in out.split("libeppscore.so")[1].split("\n")[0]


@pytest.mark.gpu
def test_host_mirror_reference_tests_on_gpu():
    exe = _pkg.load_build().build_host_test()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout
