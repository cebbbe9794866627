"""CPU-side checks of the drop-in boundary: the library builds, loads, and exports exactly the
symbols include/eppscore.h declares; without a GPU it refuses to create an engine (no CPU path)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg():
    _pkg.load_build().build()
    return _pkg.load()


def test_header_symbols_all_exported(pkg):
    header = open(os.path.join(ROOT, "include", "eppscore.h")).read()
    declared = set(re.findall(r"\b(eppscore_[a-z0-9_]+)\s*\(", header))
    bound = {name for name, _, _ in pkg.ABI_SYMBOLS}
    assert declared == bound, (declared ^ bound)
    L = pkg.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.eppscore_abi_version() == 3


def test_struct_sizes_match_header_layout(pkg):
    # natural C layout on x86-64 (computed by hand from include/eppscore.h)
    assert C.sizeof(pkg.Config) == 4 + 4 + 32 + 64 + 4 + 4 + 4 + 4 + 8 + 4 + 4 + 8 + 4 + 4 + 8 + 4 + 4 + 16 + 96
    assert C.sizeof(pkg.LatencyParams) == 4 + 4 + 12 * 8 + 8 + 4 + 4 + 5 * 8
    assert C.sizeof(pkg.Snapshot) == 16 + 7 * 8 + 4 * 8 + 8 + 8 + 4 * 8
    assert C.sizeof(pkg.Batch) == 16 + 8 + 6 * 8 + 16 + 4 * 8 + 7 * 8 + 8 + 5 * 8
    assert C.sizeof(pkg.Stats) == 8 + 8 + 8 + 8 * 8 + 8


def test_default_config_is_reference_default(pkg):
    cfg = pkg.default_config()
    # pkg/epp/config/loader/defaults.go:46-103: queue 2, kv 2, prefix 3
    assert cfg.n_scorers == 3
    assert [cfg.scorer_kind[i] for i in range(3)] == [pkg.SCORER["queue"], pkg.SCORER["kv"], pkg.SCORER["prefix"]]
    assert [cfg.scorer_weight[i] for i in range(3)] == [2.0, 2.0, 3.0]
    assert cfg.block_chars == 64 and cfg.max_blocks == 256 and cfg.lru_capacity_default == 31250
    assert cfg.token_load_threshold == 4194304.0  # token_load.go:33
    assert cfg.pick_mode == pkg.PICK_MAX_SCORE    # max-score-picker (loader/defaults.go)
    lp = pkg.latency_params()
    # predictedlatency/plugin.go:128-136 and scorer/latency/plugin.go:83-90
    assert (lp.slo_buffer_factor, lp.streaming_mode, lp.has_predictions) == (1.0, 0, 1)
    assert (lp.ttft_weight, lp.tpot_weight, lp.strategy_most) == (0.8, 0.2, 0)
    assert (lp.composite_kv, lp.composite_queue, lp.composite_prefix) == (1.0, 1.0, 1.0)


def test_host_helpers_match_oracle(pkg, xxh_kat):
    from oracle import oracle_py as o
    L = pkg.lib()
    for v in xxh_kat["raw"][:40]:
        b = bytes.fromhex(v["hex"])
        assert f"{L.eppscore_xxh64(b, len(b), 0):016x}" == v["xxh64"]
    assert pkg.Engine.model_seed("test-model1") == 0x55B9CE9184DD8509 == o.model_seed("test-model1")
    assert pkg.Engine.model_seed("m", "salt") == o.model_seed("m", "salt")


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    with pytest.raises(pkg.EppscoreError) as ei:
        pkg.Engine()
    assert ei.value.code == -5  # EPPSCORE_ERR_NO_DEVICE


def test_invalid_config_rejected(pkg):
    cfg = pkg.default_config()
    cfg.struct_size = 3
    h = C.c_void_p()
    assert pkg.lib().eppscore_create(0, C.byref(cfg), C.byref(h)) == -1
    cfg = pkg.default_config([("queue", 1.0)])
    cfg.scorer_kind[0] = 99
    assert pkg.lib().eppscore_create(0, C.byref(cfg), C.byref(h)) == -1
    assert b"scorer" in pkg.lib().eppscore_last_error(None)
    # profile validation happens before any device is touched, so it is checkable on a CPU-only box
    L = pkg.lib()

    def rejected(cfg, needle):
        assert L.eppscore_create(0, C.byref(cfg), C.byref(h)) == -1
        assert needle in L.eppscore_last_error(None), L.eppscore_last_error(None)

    rejected(pkg.default_config([("latency", 1.0), ("latency", 2.0)]), b"at most one latency scorer")
    rejected(pkg.default_config([("kv", 1.0)], pick_mode=7), b"pick_mode")
    rejected(pkg.default_config([("kv", 1.0)], filters=[(pkg.FILTER_SLO_HEADROOM_TIER, (0.01,))]), b"need the latency scorer")
    rejected(pkg.default_config([("latency", 1.0), ("queue", 1.0)], filters=[(pkg.FILTER_SLO_HEADROOM_TIER, (0.01,))]), b"queue / running")
    rejected(pkg.default_config([("latency", 1.0)], filters=[(pkg.FILTER_SLO_HEADROOM_TIER, (1.5,))]), b"epsilonExploreNeg")   # sloheadroomtier/plugin.go:66-68
    rejected(pkg.default_config([("latency", 1.0)], filters=[(pkg.FILTER_PREFIX_AFFINITY, (1.2, 0.0, 0.0))]), b"prefix-cache-affinity")  # prefixcacheaffinity/plugin.go:80-91
    rejected(pkg.default_config([("latency", 1.0)], filters=[(9, (0.0,))]), b"unknown filter kind")
    cfg = pkg.default_config([("latency", 1.0)])
    cfg.n_filters = 9
    rejected(cfg, b"n_filters")


def test_host_hash_pool_matches_oracle(pkg, xxh_kat):
    """eppscore_hash_prompts_host (the library's host-side hashPrompt, worker pool) against the chained XXH64 KATs and the
    oracle on ragged prompts, odd block sizes, truncation — no GPU involved."""
    from oracle import oracle_py as o
    from tests.helpers import synth_ragged_prompts
    for c in xxh_kat["chains"][:40]:
        p = np.frombuffer(bytes.fromhex(c["prompt_hex"]), np.uint8)
        h, n = pkg.Engine.hash_prompts_host(p if len(p) else np.zeros(1, np.uint8), np.array([0, len(p)], np.int64),
                                            np.array([int(c["seed"], 16)], np.uint64), block_chars=c["block_chars"],
                                            max_blocks=c["max_blocks"], n_threads=1)
        want = [int(x, 16) for x in c["hashes"]]
        assert int(n[0]) == len(want) and [int(x) for x in h[0, : n[0]]] == want
    data, off = synth_ragged_prompts(3000, max_len=900, seed=3)
    seeds = np.arange(3000, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    for bc, mb in ((64, 256), (32, 7), (20, 9), (4, 300)):
        h, n = pkg.Engine.hash_prompts_host(data, off, seeds, block_chars=bc, max_blocks=mb, n_threads=8)
        for r in range(0, 3000, 7):
            want = o.hash_prompt(data[off[r]: off[r + 1]], int(seeds[r]), bc, mb)
            assert int(n[r]) == len(want) and np.array_equal(h[r, : n[r]], want), (bc, mb, r)
            assert not h[r, n[r]:].any()


def test_host_hash_eight_lane_path_matches_oracle(pkg):
    """The eight-requests-per-register form of the host hashPrompt (csrc/host_hash_simd.cpp; taken when the CPU has AVX-512 F + DQ,
    the block size is a multiple of 32 and eight consecutive requests have the same number of full blocks) against the oracle:
    uniform prompts (every group takes it), uniform prompts with a trailing partial block, truncation at max_blocks, explicit
    lengths with padded starts, and a mix of uniform runs and ragged prompts so that groups switch between the two paths.
    On a CPU without AVX-512 the same calls run the scalar path: the test still holds."""
    from oracle import oracle_py as o
    rng = np.random.default_rng(5)

    def check(data, off, seeds, bc, mb, lens=None, threads=4, every=1):
        h, n = pkg.Engine.hash_prompts_host(data, off, seeds, prompt_len=lens, block_chars=bc, max_blocks=mb, n_threads=threads)
        for r in range(0, len(off) - 1, every):
            ln = int(lens[r]) if lens is not None else int(off[r + 1] - off[r])
            want = o.hash_prompt(data[off[r]: off[r] + ln], int(seeds[r]), bc, mb)
            assert int(n[r]) == len(want), (bc, mb, r, int(n[r]), len(want))
            assert np.array_equal(h[r, : n[r]], want), (bc, mb, r)
            assert not h[r, n[r]:].any()

    R = 203                                                     # not a multiple of 8: a scalar tail in every chunk
    for plen, bc, mb in ((2048, 64, 32), (2048 + 37, 64, 64), (4096, 64, 20), (640, 32, 256), (960, 96, 16), (50, 64, 8)):
        data = rng.integers(0, 256, size=R * plen, dtype=np.uint8)
        off = np.arange(R + 1, dtype=np.int64) * plen
        seeds = rng.integers(0, 2 ** 63, size=R, dtype=np.uint64)
        check(data, off, seeds, bc, mb, threads=1)
        check(data, off, seeds, bc, mb, threads=4)
    # padded starts + explicit lengths (the host layer's layout), lengths equal within runs of 8..40 requests, ragged between
    lens, starts, total = [], [], 0
    while len(lens) < 600:
        run, ln = int(rng.integers(1, 40)), int(rng.integers(0, 1500))
        for _ in range(run):
            ln_r = ln if rng.random() < 0.9 else int(rng.integers(0, 1500))
            total = (total + 15) // 16 * 16
            starts.append(total)
            lens.append(ln_r)
            total += ln_r
    Rm = len(lens)
    data = rng.integers(0, 256, size=total + 64, dtype=np.uint8)
    off = np.array(starts + [total], np.int64)
    seeds = rng.integers(0, 2 ** 63, size=Rm, dtype=np.uint64)
    check(data, off, seeds, 64, 256, lens=np.array(lens, np.int32), threads=3)
    check(data, off, seeds, 32, 11, lens=np.array(lens, np.int32), threads=2)
