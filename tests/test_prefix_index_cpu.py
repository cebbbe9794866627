"""The product's host-side prefix index (csrc/prefix_index.hpp: per-endpoint LRUs with golang-lru semantics,
open-addressing slot table, INTERNED bitset rows) against the oracle's indexer (oracle/oracle.c, the restatement
of approximateprefix/indexer.go) under random operation sequences.  CPU only: the class is plain C++."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_py as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "prefix_index_capi.cpp")
OUT = os.path.join(ROOT, "tests", "cpp", "_build", "libpit.so")


@pytest.fixture(scope="module")
def pit():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC, os.path.join(ROOT, "gateway-api-inference-extension_b200", "csrc", "prefix_index.hpp"),
            os.path.join(ROOT, "gateway-api-inference-extension_b200", "csrc", "kernels.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-D_GLIBCXX_ASSERTIONS", "-I/usr/local/cuda/include", SRC, "-o", OUT])  # bounds-checked std::vector
    L = C.CDLL(OUT)
    L.pit_new.restype = C.c_void_p
    L.pit_new.argtypes = [C.c_int, C.c_longlong, C.c_int]
    L.pit_free.argtypes = [C.c_void_p]
    L.pit_add.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.pit_apply.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int]
    L.pit_remove_endpoint.argtypes = [C.c_void_p, C.c_int]
    L.pit_lru_len.argtypes = [C.c_void_p, C.c_int]
    L.pit_lru_keys.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.pit_get.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
    for f in ("pit_n_live", "pit_n_keys", "pit_n_rows", "pit_lru_entries", "pit_dirty"):
        getattr(L, f).restype = C.c_longlong
        getattr(L, f).argtypes = [C.c_void_p]
    L.pit_clear_dirty.argtypes = [C.c_void_p]
    return L


def _get(L, h, key):
    buf = np.zeros(8192, np.int32)
    n = L.pit_get(h, int(key), buf.ctypes.data, len(buf))
    assert n >= 0, "slot count != popcount of its (interned) row"
    return set(int(x) for x in buf[:n])


@pytest.mark.parametrize("M,lru,seed", [(40, 12, 0), (300, 50, 1), (1024, 9, 2), (5000, 30, 3), (257, 20, 4), (513, 40, 5), (8192, 6, 6)])
def test_random_ops_match_oracle(pit, M, lru, seed):
    L = pit
    rng = np.random.Generator(np.random.PCG64(seed))
    h = L.pit_new(M, 1 << 14, lru)
    idx = o.Index(lru)
    universe = rng.integers(1, 2 ** 63, size=400, dtype=np.uint64)
    # chains: consecutive hashes, as prompts produce them — the case the interning is for
    chains = [universe[s:s + int(rng.integers(1, 24))] for s in rng.integers(0, 380, size=60)]
    used_eps = set()
    for step in range(1500):
        op = rng.random()
        ep = int(rng.integers(0, M))
        if op < 0.80:
            ch = chains[int(rng.integers(0, len(chains)))]
            cap = int(rng.integers(3, 40)) if rng.random() < 0.3 else 0
            assert L.pit_add(h, ch.ctypes.data, len(ch), ep, cap) == 0
            idx.add(ch, ep, cap)
            used_eps.add(ep)
        elif op < 0.85 and used_eps:
            ep = int(rng.choice(sorted(used_eps)))
            L.pit_remove_endpoint(h, ep)
            idx.remove_pod(ep)
            used_eps.discard(ep)
        if step % 100 == 99:
            for key in universe[:: 7]:
                assert _get(L, h, key) == idx.get(int(key)), (step, int(key))
            for e in list(used_eps)[:20]:
                n = idx.lru_len(e)
                assert L.pit_lru_len(h, e) == n
                out = np.zeros(max(n, 1), np.uint64)
                L.pit_lru_keys(h, e, out.ctypes.data, len(out))
                assert [int(x) for x in out[:n]] == idx.lru_keys(e)
            assert L.pit_n_live(h) == idx.num_hashes()
    for key in universe:
        assert _get(L, h, key) == idx.get(int(key))
    # interning really shares rows: far fewer rows than hashes with a non-empty set when chains repeat
    assert L.pit_n_rows(h) <= L.pit_n_keys(h) + 1
    L.pit_free(h)


def test_interning_shares_rows_and_tracks_dirty_words(pit):
    L = pit
    h = L.pit_new(1024, 1 << 12, 1000)
    chain = np.arange(100, 132, dtype=np.uint64)          # one 32-block prompt
    for ep in (3, 700, 41):
        assert L.pit_add(h, chain.ctypes.data, len(chain), ep, 0) == 0
    assert L.pit_n_keys(h) == 32 and L.pit_n_live(h) == 32
    assert L.pit_n_rows(h) <= 1 + 3 + 3                     # empty row + a handful of distinct sets, NOT 32
    for k in chain:
        assert _get(L, h, k) == {3, 41, 700}
    L.pit_clear_dirty(h)
    other = np.arange(100, 116, dtype=np.uint64)           # a second prompt sharing the first 16 blocks
    assert L.pit_add(h, other.ctypes.data, len(other), 9, 0) == 0
    assert _get(L, h, 100) == {3, 9, 41, 700} and _get(L, h, 131) == {3, 41, 700}
    assert 0 < L.pit_dirty(h) <= 16 + 40                     # 16 re-pointed slots + one new row's words
    # raw deltas (a host with its own LRU)
    assert L.pit_apply(h, 100, 9, 1) == 0
    assert _get(L, h, 100) == {3, 41, 700}
    for ep in (3, 41, 700):
        L.pit_apply(h, 131, ep, 1)
    assert _get(L, h, 131) == set() and L.pit_n_live(h) == 31
    L.pit_free(h)


def test_table_grows_on_demand(pit):
    L = pit
    h = L.pit_new(64, 16, 10000)                             # room for 16 hashes only
    idx = o.Index(10000)
    keys = np.arange(1, 2001, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    for i in range(0, len(keys), 50):
        ch = keys[i:i + 50]
        assert L.pit_add(h, ch.ctypes.data, len(ch), i % 64, 0) == 0
        idx.add(ch, i % 64, 0)
    assert L.pit_n_keys(h) == 2000
    for k in keys[::13]:
        assert _get(L, h, k) == idx.get(int(k))
    L.pit_free(h)
