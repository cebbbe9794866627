"""Host pieces above the C ABI that need no engine (gateway-api-inference-extension_b200/host/coalescer.hpp), CPU only:
* the coalescing front under ThreadSanitizer (64 caller threads, mock backend), the LoRA metric-label parser on the
  reference's vectors, top-k from a score row — tests/cpp/host_logic_test.cpp;
* the small-batch CPU route (SmallBatchCpu: product code, float64, reference operation order) against the oracle on random
  snapshots, scorer orders, negative weights and candidate subsets, incl. BASELINE config A (1 request x 4 pods, queue scorer)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_py as o
from tests.helpers import synth_snapshot

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_logic_test.cpp")
HDRS = [os.path.join(ROOT, "gateway-api-inference-extension_b200", "host", h) for h in ("coalescer.hpp", "epp_scheduler.hpp")]
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")


def _stale(out):
    return not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in [SRC] + HDRS)


def test_coalescer_lora_labels_topk_under_tsan():
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "host_logic_tsan")
    if _stale(exe):
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fsanitize=thread", "-pthread", SRC, "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ThreadSanitizer" not in r.stderr, r.stderr
    assert "host logic: all ok" in r.stdout


@pytest.fixture(scope="module")
def sbc():
    os.makedirs(BUILD, exist_ok=True)
    lib = os.path.join(BUILD, "libhostlogic.so")
    if _stale(lib):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", SRC, "-o", lib])
    L = C.CDLL(lib)
    L.sbc_schedule.restype = C.c_int
    L.sbc_schedule.argtypes = [C.c_int] + [C.c_void_p] * 8 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def _cpu(L, sd, tokens, scorers, adapter, mask):
    M = len(sd["kv_usage"])
    kinds = np.array([k for k, _ in scorers], np.int32)
    weights = np.array([w for _, w in scorers], np.float64)
    pick, score, ties = C.c_int(), C.c_double(), C.c_int()
    arrs = [np.ascontiguousarray(sd["kv_usage"], np.float64), np.ascontiguousarray(sd["queue"], np.int64), np.ascontiguousarray(sd["running"], np.int64),
            np.ascontiguousarray(sd["lora_active"], np.uint64), np.ascontiguousarray(sd["lora_waiting"], np.uint64),
            np.ascontiguousarray(sd["lora_nmodels"], np.int32), np.ascontiguousarray(sd["lora_max"], np.int32)]
    tok = np.ascontiguousarray(tokens, np.int64) if tokens is not None else None
    msk = np.ascontiguousarray(mask, np.uint32) if mask is not None else None
    rc = L.sbc_schedule(M, *[a.ctypes.data for a in arrs], tok.ctypes.data if tok is not None else None, len(kinds), kinds.ctypes.data,
                        weights.ctypes.data, adapter, msk.ctypes.data if msk is not None else None, C.byref(pick), C.byref(score), C.byref(ties))
    assert rc == 0
    return pick.value, score.value, ties.value


@pytest.mark.parametrize("M,seed", [(4, 0), (7, 1), (64, 2), (300, 3), (1024, 4)])
def test_small_batch_cpu_route_matches_oracle(sbc, M, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = synth_snapshot(M, seed=seed, tie_heavy=(seed % 2 == 0))
    tokens = rng.integers(0, 6_000_000, M).astype(np.int64)
    snap = o.SnapshotData(**sd, inflight_tokens=tokens)
    pool = [(o.SCORER_QUEUE, 2.0), (o.SCORER_KV_CACHE, 2.0), (o.SCORER_LORA, 1.0), (o.SCORER_RUNNING, 0.5), (o.SCORER_TOKEN_LOAD, 1.5),
            (o.SCORER_KV_CACHE, -1.0)]
    for trial in range(60):
        scorers = [pool[i] for i in rng.permutation(len(pool))[: int(rng.integers(1, 6))]]
        prof = o.make_profile(scorers)
        adapter = int(rng.integers(-1, 64))
        mask = None
        if trial % 3 == 1:
            keep = rng.random(M) < 0.5
            keep[int(rng.integers(0, M))] = True
            mask = np.zeros((M + 31) // 32, np.uint32)
            for m in np.nonzero(keep)[0]:
                mask[m >> 5] |= np.uint32(1 << (m & 31))
        want = o.schedule_batch(snap, prof, None, 1, adapter_id=np.array([adapter], np.int32),
                                cand_mask=mask.reshape(1, -1) if mask is not None else None)
        pick, score, ties = _cpu(sbc, sd, tokens, scorers, adapter, mask)
        assert (pick, ties) == (int(want["pick"][0]), int(want["tie_count"][0])), (trial, scorers)
        assert np.float64(score).view(np.uint64) == want["pick_score"][0].view(np.uint64), (trial, scorers)


def test_config_A_one_request_four_pods_queue_scorer(sbc):
    """BASELINE.json config A through PRODUCT code: queue_test.go:35-69's table."""
    for queue, want_pick, want_score, want_ties in (([10, 5, 0, 7], 2, 1.0, 1), ([5, 5, 5, 5], 0, 1.0, 4), ([0, 0, 3, 9], 0, 1.0, 2)):
        sd = dict(kv_usage=np.zeros(4), queue=np.array(queue, np.int64), running=np.zeros(4, np.int64), lora_active=np.zeros((4, 1), np.uint64),
                  lora_waiting=np.zeros((4, 1), np.uint64), lora_nmodels=np.zeros(4, np.int32), lora_max=np.zeros(4, np.int32))
        assert _cpu(sbc, sd, None, [(o.SCORER_QUEUE, 1.0)], -1, None) == (want_pick, want_score, want_ties)
