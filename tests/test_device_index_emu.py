"""The CTA program of the device-resident prefix index (csrc/prefix_table.cuh: log-structured per-endpoint LRUs with exact
golang-lru semantics + the slot table with inline sets / overflow bitset rows) instantiated with a SEQUENTIAL execution
policy (tests/cpp/index_emu.cpp) and checked against the oracle's indexer (oracle/oracle.c, the restatement of
approximateprefix/indexer.go) under random operation sequences.  CPU only; the same source runs as CUDA kernels
(prefix_index.cu) and is checked on the GPU by tests/test_gpu_parity.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_py as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "index_emu.cpp")
OUT = os.path.join(ROOT, "tests", "cpp", "_build", "libindex_emu.so")


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC, os.path.join(ROOT, "gateway-api-inference-extension_b200", "csrc", "prefix_table.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-D_GLIBCXX_ASSERTIONS", SRC, "-o", OUT])
    L = C.CDLL(OUT)
    L.emu_new.restype = C.c_void_p
    L.emu_new.argtypes = [C.c_int, C.c_longlong, C.c_int, C.c_int]
    L.emu_free.argtypes = [C.c_void_p]
    L.emu_commit.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.emu_apply.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int]
    L.emu_remove_endpoint.argtypes = [C.c_void_p, C.c_int]
    L.emu_lru_len.argtypes = [C.c_void_p, C.c_int]
    L.emu_lru_keys.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.emu_get.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
    L.emu_stat.restype = C.c_longlong
    L.emu_stat.argtypes = [C.c_void_p, C.c_int]
    return L


def _get(L, h, key):
    buf = np.zeros(8192, np.int32)
    n = L.emu_get(h, int(key), buf.ctypes.data, len(buf))
    assert n >= 0, "slot count != size of its set"
    return set(int(x) for x in buf[:n])


def _add(L, h, hashes, ep, cap=0):
    hs = np.ascontiguousarray(hashes, np.uint64)
    pick = np.array([ep], np.int32)
    nh = np.array([len(hs)], np.uint16)
    assert L.emu_commit(h, 1, pick.ctypes.data, hs.ctypes.data, nh.ctypes.data, max(len(hs), 1), None, cap) == 0


def _check_lrus(L, h, idx, eps):
    for e in eps:
        n = idx.lru_len(e)
        assert L.emu_lru_len(h, e) == n, e
        if n < 0:
            continue
        out = np.zeros(max(n, 1), np.uint64)
        assert L.emu_lru_keys(h, e, out.ctypes.data, len(out)) == n
        assert [int(x) for x in out[:n]] == idx.lru_keys(e), e


@pytest.mark.parametrize("M,lru,seed", [(40, 12, 0), (300, 50, 1), (1024, 9, 2), (5000, 30, 3), (257, 20, 4), (513, 40, 5), (8192, 6, 6)])
def test_random_single_adds_match_oracle(emu, M, lru, seed):
    """indexer.Add one call at a time, capacities from 3 (shorter than most calls: the strictly sequential path with its
    stale-entry quirk) to 40, RemovePod, duplicate hashes inside a call."""
    L = emu
    rng = np.random.Generator(np.random.PCG64(seed))
    h = L.emu_new(M, 1 << 10, lru, 40)
    idx = o.Index(lru)
    universe = rng.integers(1, 2 ** 63, size=400, dtype=np.uint64)
    chains = [universe[s:s + int(rng.integers(1, 24))] for s in rng.integers(0, 380, size=60)]
    chains.append(np.concatenate([universe[5:9], universe[5:7], universe[8:9]]))  # duplicates inside one call
    used = set()
    for step in range(1500):
        op = rng.random()
        ep = int(rng.integers(0, M))
        if op < 0.80:
            ch = chains[int(rng.integers(0, len(chains)))]
            cap = int(rng.integers(3, 40)) if rng.random() < 0.3 else 0
            _add(L, h, ch, ep, cap)
            idx.add(ch, ep, cap)
            used.add(ep)
        elif op < 0.85 and used:
            ep = int(rng.choice(sorted(used)))
            L.emu_remove_endpoint(h, ep)
            idx.remove_pod(ep)
            used.discard(ep)
        if step % 100 == 99:
            for key in universe[::7]:
                assert _get(L, h, key) == idx.get(int(key)), (step, int(key))
            _check_lrus(L, h, idx, list(used)[:20])
            assert L.emu_stat(h, 0) == idx.num_hashes()
            assert L.emu_stat(h, 7) == 0                  # every slot canonical (sorted members, 0xFFFF filler)
    for key in universe:
        assert _get(L, h, key) == idx.get(int(key))
    _check_lrus(L, h, idx, sorted(used))
    L.emu_free(h)


@pytest.mark.parametrize("M,lru,R,nmax,seed", [(64, 31250, 4000, 32, 0), (16, 700, 3000, 40, 1), (8, 50, 2000, 64, 2),
                                                (4, 2000, 6000, 300, 3), (32, 5, 500, 12, 4), (3, 1500, 5000, 32, 5)])
def test_batch_commit_matches_oracle(emu, M, lru, R, nmax, seed):
    """PreRequest for whole batches: many requests per endpoint (several parallel chunks per CTA, chunks that split calls,
    evictions reaching back over earlier batches, log compaction / map rebuilds), shared prefixes (re-touched keys),
    per-endpoint CacheNumBlocks, requests without a pick."""
    L = emu
    rng = np.random.Generator(np.random.PCG64(100 + seed))
    h = L.emu_new(M, 1 << 8, lru, max(lru, 3000))
    idx = o.Index(lru)
    groups = [rng.integers(1, 2 ** 63, size=nmax, dtype=np.uint64) for _ in range(12)]
    caps = rng.integers(max(lru // 2, 3), max(lru, 3000), size=M).astype(np.int32) if seed % 2 else None
    seen = set()
    for batch in range(4):
        hashes = np.zeros((R, nmax), np.uint64)
        nh = np.zeros(R, np.uint16)
        pick = rng.integers(-1, M, size=R).astype(np.int32)
        for r in range(R):
            n = int(rng.integers(0, nmax + 1))
            g = groups[int(rng.integers(0, len(groups)))]
            k = int(rng.integers(0, n + 1))
            hashes[r, :k] = g[:k]                                          # shared prefix
            hashes[r, k:n] = rng.integers(1, 2 ** 63, size=n - k, dtype=np.uint64)  # unique tail
            nh[r] = n
        capp = caps.ctypes.data if caps is not None else None
        assert L.emu_commit(h, R, pick.ctypes.data, hashes.ctypes.data, nh.ctypes.data, nmax, capp, 0) == 0
        idx.commit(pick, hashes, nh, caps)
        seen.update(int(p) for p in pick if p >= 0)
        _check_lrus(L, h, idx, sorted(seen))
        assert L.emu_stat(h, 0) == idx.num_hashes()
        assert L.emu_stat(h, 7) == 0
        for g in groups:
            for key in g[::5]:
                assert _get(L, h, key) == idx.get(int(key))
        for r in rng.integers(0, R, size=60):
            for key in hashes[r, : nh[r]][::3]:
                assert _get(L, h, key) == idx.get(int(key))
    assert L.emu_stat(h, 5) == sum(max(idx.lru_len(e), 0) for e in range(M))
    L.emu_free(h)


def test_inline_sets_overflow_rows_and_raw_deltas(emu):
    L = emu
    M = 1024
    h = L.emu_new(M, 1 << 12, 1000, 0)
    idx = o.Index(1000)
    chain = np.arange(100, 132, dtype=np.uint64)          # one 32-block prompt cached on more and more endpoints
    eps = [3, 700, 41, 9, 1023, 512, 77, 300, 2, 1000]
    for ep in eps:
        _add(L, h, chain, ep)
        idx.add(chain, ep, 0)
    assert L.emu_stat(h, 3) == 0                          # ten members still fit the slot
    for k in chain:
        assert _get(L, h, k) == set(eps)
    for ep in (5, 6, 7, 900):                             # the eleventh member moves each set to a bitset row
        _add(L, h, chain[:16], ep)
        idx.add(chain[:16], ep, 0)
    assert L.emu_stat(h, 3) == 16
    assert _get(L, h, 100) == set(eps) | {5, 6, 7, 900} and _get(L, h, 131) == set(eps)
    # raw deltas (a host with its own LRU): evictions empty a set -> its row returns to the pool, the key reads as absent
    for ep in sorted(set(eps) | {5, 6, 7, 900}):
        assert L.emu_apply(h, 100, ep, 1) == 0
    assert _get(L, h, 100) == set() and L.emu_stat(h, 3) == 15 and L.emu_stat(h, 0) == 31
    assert L.emu_apply(h, 100, 44, 0) == 0
    assert _get(L, h, 100) == {44} and L.emu_stat(h, 0) == 32
    assert L.emu_stat(h, 7) == 0                          # the slot that was a row is a canonical inline set again
    L.emu_free(h)


def test_universal_prefix_on_every_endpoint(emu):
    """A system prompt cached everywhere: sets of all M endpoints, then every endpoint removed again."""
    L = emu
    M = 300
    h = L.emu_new(M, 1 << 8, 100, 0)
    idx = o.Index(100)
    sysp = np.arange(1, 9, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    for ep in range(M):
        _add(L, h, sysp, ep)
        idx.add(sysp, ep, 0)
    for k in sysp:
        assert _get(L, h, k) == set(range(M))
    for ep in range(0, M, 2):
        L.emu_remove_endpoint(h, ep)
        idx.remove_pod(ep)
    for k in sysp:
        assert _get(L, h, k) == set(range(1, M, 2))
    assert L.emu_lru_len(h, 0) == -1 and L.emu_lru_len(h, 1) == 8
    L.emu_free(h)


def test_churn_keeps_the_table_bounded(emu):
    """LRU churn: emptied keys must be reclaimed — the table may not grow without bound while only a
    few hundred hashes are live."""
    L = emu
    M, lru = 4, 100
    h = L.emu_new(M, 1 << 9, lru, 0)
    idx = o.Index(lru)
    rng = np.random.Generator(np.random.PCG64(9))
    for it in range(3000):
        ch = rng.integers(1, 2 ** 63, size=64, dtype=np.uint64)
        ep = it % M
        _add(L, h, ch, ep)
        idx.add(ch, ep, 0)
    assert L.emu_stat(h, 0) == idx.num_hashes() == M * lru
    assert L.emu_stat(h, 4) > 5                              # rebuilt many times ...
    assert L.emu_stat(h, 2) <= 4096                          # ... and never grew: 400 live hashes, 192K hashes seen
    assert L.emu_stat(h, 1) <= L.emu_stat(h, 2) // 2
    _check_lrus(L, h, idx, range(M))
    assert L.emu_stat(h, 6) <= M * (4 * lru + 2 * 1024 + 64)  # the logs stay inside their rings
    L.emu_free(h)
