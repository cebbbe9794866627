"""The C++ host mirror's own test program (gateway-api-inference-extension_b200/host/host_test.cpp: the reference's
scheduler tests restated — TestSchedule, the filter chain, the integration routing scenarios, TestPrefixPluginCompletion via
PreRequest, the token-load / latency-scorer / picker tests, top-k, and 32 caller threads through the coalescing front) on the
CPU: the program is linked against tests/cpp/oracle_backed_abi.cpp, which implements the ten C-ABI entry points the host layer
calls on top of the oracle.  What this exercises is the HOST code (Scheduler::ScheduleBatch, PackSnapshot, the adapter
dictionary, filter masks, prompt packing, result hand-back, PreRequest, the small-batch host route beside the engine route);
`tests/test_host_cpp.py` (-m gpu) runs the same program against libeppscore.so on a B200."""
import os
import subprocess

from oracle import oracle_py as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "gateway-api-inference-extension_b200", "host")
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")
SRCS = [os.path.join(HOST, "host_test.cpp"), os.path.join(ROOT, "tests", "cpp", "oracle_backed_abi.cpp")]
DEPS = SRCS + [os.path.join(HOST, h) for h in ("epp_scheduler.hpp", "epp_types.hpp", "host_eval.hpp", "coalescer.hpp")] + \
    [os.path.join(ROOT, "include", "eppscore.h"), os.path.join(ROOT, "oracle", "oracle.h")]


def _build(name, extra, srcs=None):
    global SRCS
    srcs = srcs or SRCS
    o.build()                                    # oracle/_build/liboracle.so (gcc)
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, name)
    libdir = os.path.join(ROOT, "oracle", "_build")
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in DEPS + srcs + [os.path.join(libdir, "liboracle.so")]):
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-pthread"] + extra + srcs +
                              ["-I" + os.path.join(ROOT, "include"), "-L" + libdir, "-loracle", "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def test_host_mirror_reference_tests_on_the_oracle_backed_abi():
    exe = _build("host_test_cpu", [])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "pass 0 (CpuBatchThreshold = 0) done" in r.stdout and "pass 1 (CpuBatchThreshold = 8) done" in r.stdout
    assert "coalesced schedule: 640 requests" in r.stdout and "all checks passed" in r.stdout


def test_stable_endpoint_ids_across_reordered_extended_and_reduced_lists():
    """tests/cpp/host_ids_test.cpp: prefix affinity follows the endpoint's NAME (ServerID) across candidate lists of different
    order and membership, RemovePod drops the history and recycles the id, top-k reports positions of the current list."""
    exe = _build("host_ids_test", [], [os.path.join(ROOT, "tests", "cpp", "host_ids_test.cpp"), os.path.join(ROOT, "tests", "cpp", "oracle_backed_abi.cpp")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host ids: all checks passed" in r.stdout


def test_host_mirror_under_thread_sanitizer():
    """The same program under ThreadSanitizer: the coalescing front drives the REAL Scheduler class from 32 caller threads
    (one engine call at a time is the engine's contract — the front has to provide it)."""
    exe = _build("host_test_cpu_tsan", ["-fsanitize=thread"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ThreadSanitizer" not in r.stderr, r.stderr
    assert "all checks passed" in r.stdout
