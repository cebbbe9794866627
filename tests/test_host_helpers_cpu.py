"""CPU-side checks of the C++ host mirror's pure helpers (host/epp_scheduler.hpp) against the oracle."""
import ctypes as C
import os
import random
import subprocess

import pytest

from oracle import oracle_py as o
from tests.test_oracle_golden import FIELDS_CASES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_helpers_capi.cpp")
OUT = os.path.join(ROOT, "tests", "cpp", "_build", "libhosthelpers.so")


@pytest.fixture(scope="module")
def hh():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC, os.path.join(ROOT, "gateway-api-inference-extension_b200", "host", "epp_scheduler.hpp"),
            os.path.join(ROOT, "include", "eppscore.h")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", SRC, "-o", OUT])
    L = C.CDLL(OUT)
    L.epp_count_fields.restype = C.c_int
    L.epp_count_fields.argtypes = [C.c_char_p, C.c_int]
    return L


def test_count_fields_matches_go_semantics(hh):
    """CountFields == len(strings.Fields(s)) (predictedlatency/plugin.go:286) as restated by the oracle's rune decoder."""
    for s, want in FIELDS_CASES:
        assert hh.epp_count_fields(s, len(s)) == want, s
    rnd = random.Random(3)
    alphabet = [0x20, 0x09, 0x41, 0xC2, 0x85, 0xA0, 0xE1, 0x9A, 0x80, 0xE2, 0x81, 0x9F, 0xE3, 0x8A, 0xA8, 0xA9, 0xAF, 0xF0, 0xF4,
                0xED, 0xE0, 0xC0, 0xFF, 0x0A, 0x8B, 0x90, 0xBF, 0x00]
    for _ in range(50000):
        s = bytes(rnd.choice(alphabet) for _ in range(rnd.randint(0, 40)))
        assert hh.epp_count_fields(s, len(s)) == o.count_fields(s), s
