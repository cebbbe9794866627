"""ctypes binding of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs — never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

MAX_SCORERS = 8
SCORER_QUEUE, SCORER_KV_CACHE, SCORER_PREFIX, SCORER_LORA, SCORER_RUNNING = 0, 1, 2, 3, 4
SCORER_LATENCY, SCORER_TOKEN_LOAD = 5, 6
SCORER_ENDPOINT_COL0 = 8
SCORER_PAIR_COL0 = 16
TIE_LOWEST_INDEX, TIE_SEEDED_RANDOM = 0, 1
PICK_MAX_SCORE, PICK_WEIGHTED_RANDOM, PICK_RANDOM = 0, 1, 2
FILTER_PREFIX_AFFINITY, FILTER_SLO_HEADROOM_TIER = 1, 2


_GOSHAPE_PATH = os.path.join(_HERE, "_build", "libgoshape.so")


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle.h", "Makefile", "goshape.cpp")]
    stale = (not os.path.exists(_LIB_PATH)) or (not os.path.exists(_GOSHAPE_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _LIB_PATH


class LatencyParams(C.Structure):
    _fields_ = [("ttft_intercept", C.c_double), ("ttft_kv", C.c_double), ("ttft_input", C.c_double),
                ("ttft_waiting", C.c_double), ("ttft_running", C.c_double), ("ttft_prefix", C.c_double),
                ("tpot_intercept", C.c_double), ("tpot_kv", C.c_double), ("tpot_input", C.c_double),
                ("tpot_waiting", C.c_double), ("tpot_running", C.c_double), ("tpot_generated", C.c_double),
                ("slo_buffer_factor", C.c_double), ("streaming_mode", C.c_int32), ("has_predictions", C.c_int32),
                ("ttft_weight", C.c_double), ("tpot_weight", C.c_double), ("strategy_most", C.c_int32),
                ("reserved", C.c_int32), ("composite_kv", C.c_double), ("composite_queue", C.c_double),
                ("composite_prefix", C.c_double)]


LATENCY_FIELDS = [f for f, _ in LatencyParams._fields_ if f != "reserved"]


def make_latency_params(**kw) -> LatencyParams:
    """Defaults: predictedlatency DefaultConfig (plugin.go:128-136) + latency-scorer DefaultConfig (plugin.go:83-90)."""
    lp = LatencyParams()
    lp.slo_buffer_factor = 1.0
    lp.streaming_mode = 0
    lp.has_predictions = 1
    lp.ttft_weight, lp.tpot_weight = 0.8, 0.2
    lp.strategy_most = 0
    lp.composite_kv = lp.composite_queue = lp.composite_prefix = 1.0
    for k, v in kw.items():
        if k not in LATENCY_FIELDS:
            raise KeyError(k)
        setattr(lp, k, v)
    return lp


class LatencyRequest(C.Structure):
    _fields_ = [("input_tokens", C.c_int64), ("ttft_slo", C.c_double), ("tpot_slo", C.c_double)]


class Profile(C.Structure):
    _fields_ = [("n_scorers", C.c_int32), ("scorer_kind", C.c_int32 * MAX_SCORERS),
                ("scorer_weight", C.c_double * MAX_SCORERS), ("tie_mode", C.c_int32),
                ("tie_seed", C.c_uint64), ("latency", C.POINTER(LatencyParams)),
                ("token_load_threshold", C.c_double), ("pick_mode", C.c_int32), ("n_filters", C.c_int32),
                ("filter_kind", C.c_int32 * 4), ("filter_param", (C.c_double * 3) * 4)]


class Snapshot(C.Structure):
    _fields_ = [("M", C.c_int32), ("lora_words", C.c_int32), ("kv_usage", C.c_void_p),
                ("queue", C.c_void_p), ("running", C.c_void_p), ("lora_active", C.c_void_p),
                ("lora_waiting", C.c_void_p), ("lora_nmodels", C.c_void_p), ("lora_max", C.c_void_p),
                ("endpoint_col", C.c_void_p * 4), ("min_tpot_slo", C.c_void_p), ("dispatched", C.c_void_p),
                ("prefill_role", C.c_void_p), ("inflight_tokens", C.c_void_p)]


class Batch(C.Structure):
    _fields_ = [("R", C.c_int32), ("request_base", C.c_int64), ("prompt_bytes", C.c_void_p),
                ("prompt_off", C.c_void_p), ("model_seed", C.c_void_p), ("hashes_in", C.c_void_p),
                ("n_hashes_in", C.c_void_p), ("hash_stride", C.c_int32), ("adapter_id", C.c_void_p),
                ("cand_mask", C.c_void_p), ("dense_feat", C.c_void_p), ("dense_total", C.c_void_p),
                ("block_chars", C.c_int32), ("max_blocks", C.c_int32), ("pick", C.c_void_p),
                ("pick_score", C.c_void_p), ("tie_count", C.c_void_p), ("tie_set", C.c_void_p),
                ("match_blocks", C.c_void_p), ("total_blocks", C.c_void_p), ("hashes_out", C.c_void_p),
                ("weighted_out", C.c_void_p), ("input_tokens", C.c_void_p), ("ttft_slo", C.c_void_p),
                ("tpot_slo", C.c_void_p), ("pred_out", C.c_void_p), ("filter_mask_out", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orc_xxh64.restype = C.c_uint64
        L.orc_xxh64.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
        L.orc_model_seed.restype = C.c_uint64
        L.orc_model_seed.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.orc_hash_prompt.restype = C.c_int32
        L.orc_hash_prompt.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]
        L.orc_index_new.restype = C.c_void_p
        L.orc_index_new.argtypes = [C.c_int32]
        L.orc_index_free.argtypes = [C.c_void_p]
        L.orc_index_add.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        L.orc_index_get.restype = C.c_int32
        L.orc_index_get.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int32]
        L.orc_index_remove_pod.argtypes = [C.c_void_p, C.c_int32]
        L.orc_index_lru_len.restype = C.c_int32
        L.orc_index_lru_len.argtypes = [C.c_void_p, C.c_int32]
        L.orc_index_lru_keys.restype = C.c_int32
        L.orc_index_lru_keys.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        L.orc_index_num_hashes.restype = C.c_int64
        L.orc_index_num_hashes.argtypes = [C.c_void_p]
        L.orc_index_pods.restype = C.c_int32
        L.orc_index_pods.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.orc_index_dump.restype = C.c_int64
        L.orc_index_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_match_longest_prefix.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        for f in ("orc_score_kv", "orc_score_queue", "orc_score_running"):
            getattr(L, f).argtypes = [C.POINTER(Snapshot), C.c_void_p, C.c_void_p]
        L.orc_score_lora.argtypes = [C.POINTER(Snapshot), C.c_void_p, C.c_int32, C.c_void_p]
        L.orc_score_prefix.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_enforce_score_range.restype = C.c_double
        L.orc_enforce_score_range.argtypes = [C.c_double]
        L.orc_tie_priority.restype = C.c_uint32
        L.orc_tie_priority.argtypes = [C.c_uint64, C.c_int64, C.c_int32]
        L.orc_schedule_one.restype = C.c_int32
        L.orc_schedule_one.argtypes = [C.POINTER(Snapshot), C.POINTER(Profile), C.c_int64, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_int32),
                                       C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p, C.c_void_p]
        L.orc_schedule_one_lat.restype = C.c_int32
        L.orc_schedule_one_lat.argtypes = [C.POINTER(Snapshot), C.POINTER(Profile), C.c_int64, C.c_int32, C.c_void_p,
                                           C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(LatencyRequest),
                                           C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int32),
                                           C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_score_token_load.argtypes = [C.POINTER(Snapshot), C.c_void_p, C.c_double, C.c_void_p]
        L.orc_latency_predict.argtypes = [C.POINTER(LatencyParams), C.c_double, C.c_int64, C.c_int64, C.c_int64,
                                          C.c_int64, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_latency_validate.argtypes = [C.POINTER(LatencyParams), C.c_double, C.c_double, C.c_double, C.c_double,
                                           C.c_double, C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_score_latency_info.argtypes = [C.POINTER(LatencyParams), C.POINTER(Snapshot), C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.orc_score_latency.argtypes = [C.POINTER(LatencyParams), C.POINTER(Snapshot), C.c_void_p, C.c_void_p,
                                        C.c_int32, C.POINTER(LatencyRequest), C.c_void_p, C.c_void_p]
        L.orc_count_fields.restype = C.c_int32
        L.orc_count_fields.argtypes = [C.c_char_p, C.c_int64]
        L.orc_uniform01.restype = C.c_double
        L.orc_uniform01.argtypes = [C.c_uint64, C.c_int64, C.c_int32]
        L.orc_neg_log.restype = C.c_double
        L.orc_neg_log.argtypes = [C.c_double]
        L.orc_schedule_batch.restype = C.c_int32
        L.orc_schedule_batch.argtypes = [C.POINTER(Snapshot), C.POINTER(Profile), C.c_void_p, C.POINTER(Batch), C.c_int32]
        L.orc_commit_picks.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        _lib = L
    return _lib


_goshape = None


def goshape_lib():
    """The "Go-shape" CPU baseline (oracle/goshape.cpp): maps, per-request clones, shuffle + stable sort."""
    global _goshape
    if _goshape is None:
        lib()
        G = C.CDLL(_GOSHAPE_PATH)
        G.orc_goshape_schedule_batch.restype = C.c_int32
        G.orc_goshape_schedule_batch.argtypes = [C.POINTER(Snapshot), C.POINTER(Profile), C.c_void_p, C.POINTER(Batch), C.c_int32,
                                                 C.c_uint64]
        _goshape = G
    return _goshape


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _arr(a, dtype):
    if a is None:
        return None
    return np.ascontiguousarray(a, dtype=dtype)


def xxh64(data: bytes, seed: int = 0) -> int:
    return lib().orc_xxh64(data, len(data), seed)


def model_seed(model: bytes | str, salt: bytes | str = b"") -> int:
    if isinstance(model, str):
        model = model.encode()
    if isinstance(salt, str):
        salt = salt.encode()
    return lib().orc_model_seed(model, len(model), salt, len(salt))


def hash_prompt(prompt: bytes, seed: int, block_chars: int, max_blocks: int) -> np.ndarray:
    cap = max(1, min(max_blocks, len(prompt) // max(block_chars, 1) + 1)) if block_chars > 0 else 1
    out = np.zeros(cap, dtype=np.uint64)
    buf = np.frombuffer(prompt, dtype=np.uint8) if len(prompt) else np.zeros(1, dtype=np.uint8)
    n = lib().orc_hash_prompt(_ptr(buf), len(prompt), seed, block_chars, max_blocks, _ptr(out), cap)
    assert n >= 0
    return out[:n].copy()


def make_profile(scorers, tie_mode=TIE_LOWEST_INDEX, tie_seed=0, latency: LatencyParams | None = None,
                 token_load_threshold: float = 0.0, pick_mode: int = 0, filters=()) -> Profile:
    """scorers: list of (kind, weight) in profile order."""
    p = Profile()
    if latency is not None:
        p._latency_keep = latency  # keep the pointee alive with the struct
        p.latency = C.pointer(latency)
    p.token_load_threshold = token_load_threshold
    p.pick_mode = pick_mode
    p.n_filters = len(filters)  # [(kind, (param0, param1, param2))]
    for i, (k, par) in enumerate(filters):
        p.filter_kind[i] = int(k)
        for j2, v in enumerate(par):
            p.filter_param[i][j2] = float(v)
    p.n_scorers = len(scorers)
    for i, (k, w) in enumerate(scorers):
        p.scorer_kind[i] = int(k)
        p.scorer_weight[i] = float(w)
    p.tie_mode = tie_mode
    p.tie_seed = tie_seed
    return p


class SnapshotData:
    """Owns the numpy arrays a Snapshot struct points at."""

    def __init__(self, kv_usage, queue, running=None, lora_active=None, lora_waiting=None,
                 lora_nmodels=None, lora_max=None, lora_words=None, endpoint_cols=(), min_tpot_slo=None,
                 dispatched=None, prefill_role=None, inflight_tokens=None):
        self.M = len(kv_usage)
        M = self.M
        self.kv_usage = _arr(kv_usage, np.float64)
        self.queue = _arr(queue, np.int64)
        self.running = _arr(running if running is not None else np.zeros(M), np.int64)
        if lora_active is None:
            lora_words = lora_words or 1
            lora_active = np.zeros((M, lora_words), dtype=np.uint64)
            lora_waiting = np.zeros((M, lora_words), dtype=np.uint64)
        lw = lora_words or (np.asarray(lora_active).size // M if M else 1) or 1
        self.lora_active = _arr(lora_active, np.uint64).reshape(M, lw)
        self.lora_waiting = _arr(lora_waiting, np.uint64).reshape(M, lw)
        self.lora_words = lw
        self.lora_nmodels = _arr(lora_nmodels if lora_nmodels is not None else np.zeros(M), np.int32)
        self.lora_max = _arr(lora_max if lora_max is not None else np.zeros(M), np.int32)
        self.endpoint_cols = [_arr(c, np.float64) for c in endpoint_cols]
        s = Snapshot()
        s.M = M
        s.lora_words = self.lora_words
        s.kv_usage = _ptr(self.kv_usage)
        s.queue = _ptr(self.queue)
        s.running = _ptr(self.running)
        s.lora_active = _ptr(self.lora_active)
        s.lora_waiting = _ptr(self.lora_waiting)
        s.lora_nmodels = _ptr(self.lora_nmodels)
        s.lora_max = _ptr(self.lora_max)
        for i, c in enumerate(self.endpoint_cols):
            s.endpoint_col[i] = _ptr(c)
        self.min_tpot_slo = _arr(min_tpot_slo, np.float64)
        self.dispatched = _arr(dispatched, np.int32)
        self.prefill_role = _arr(prefill_role, np.uint8)
        self.inflight_tokens = _arr(inflight_tokens, np.int64)
        s.min_tpot_slo = _ptr(self.min_tpot_slo)
        s.dispatched = _ptr(self.dispatched)
        s.prefill_role = _ptr(self.prefill_role)
        s.inflight_tokens = _ptr(self.inflight_tokens)
        self.struct = s


class Index:
    def __init__(self, default_lru: int = 31250):
        self._h = lib().orc_index_new(default_lru)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_index_free(self._h)
            self._h = None

    def add(self, hashes, server: int, num_gpu_blocks: int = 0):
        h = _arr(hashes, np.uint64)
        lib().orc_index_add(self._h, _ptr(h), len(h), server, num_gpu_blocks)

    def get(self, hash_: int) -> set:
        out = np.zeros(4096, dtype=np.int32)
        n = lib().orc_index_get(self._h, int(hash_), _ptr(out), len(out))
        if n > len(out):
            out = np.zeros(n, dtype=np.int32)
            n = lib().orc_index_get(self._h, int(hash_), _ptr(out), len(out))
        return set(int(x) for x in out[:n])

    def remove_pod(self, server: int):
        lib().orc_index_remove_pod(self._h, server)

    def lru_len(self, server: int) -> int:
        return lib().orc_index_lru_len(self._h, server)

    def lru_keys(self, server: int):
        n = self.lru_len(server)
        if n < 0:
            return None
        out = np.zeros(max(n, 1), dtype=np.uint64)
        lib().orc_index_lru_keys(self._h, server, _ptr(out), len(out))
        return [int(x) for x in out[:n]]

    def num_hashes(self) -> int:
        return lib().orc_index_num_hashes(self._h)

    def pods(self):
        out = np.zeros(1 << 16, dtype=np.int32)
        n = lib().orc_index_pods(self._h, _ptr(out), len(out))
        return [int(x) for x in out[:n]]

    def dump(self):
        n = lib().orc_index_dump(self._h, None, None, 0)
        h = np.zeros(max(n, 1), dtype=np.uint64)
        s = np.zeros(max(n, 1), dtype=np.int32)
        lib().orc_index_dump(self._h, _ptr(h), _ptr(s), n)
        return h[:n], s[:n]

    def match(self, hashes, M: int) -> np.ndarray:
        h = _arr(hashes, np.uint64)
        out = np.zeros(max(M, 1), dtype=np.uint16)
        lib().orc_match_longest_prefix(self._h, _ptr(h), len(h), M, _ptr(out))
        return out[:M]

    def commit(self, pick, hashes, n_hashes, gpu_blocks=None):
        pick = _arr(pick, np.int32)
        hashes = _arr(hashes, np.uint64)
        n_hashes = _arr(n_hashes, np.uint16)
        gb = _arr(gpu_blocks, np.int32)
        stride = hashes.shape[1] if hashes.ndim == 2 else 0
        lib().orc_commit_picks(self._h, len(pick), _ptr(pick), _ptr(hashes), _ptr(n_hashes), stride, _ptr(gb))


def count_fields(data: bytes) -> int:
    return lib().orc_count_fields(data, len(data))


def latency_predict(lp: LatencyParams, kv, input_tokens, waiting, running, prefix_score, generated=1):
    t, p = C.c_double(), C.c_double()
    lib().orc_latency_predict(C.byref(lp), kv, input_tokens, waiting, running, generated, prefix_score, C.byref(t),
                              C.byref(p))
    return t.value, p.value


def latency_validate(lp: LatencyParams, ttft, tpot, ttft_slo, tpot_slo, pod_min_tpot_slo=0.0, neutralize=False):
    ok = np.zeros(3, np.int32)
    hr = np.zeros(2, np.float64)
    lib().orc_latency_validate(C.byref(lp), ttft, tpot, ttft_slo, tpot_slo, pod_min_tpot_slo, 1 if neutralize else 0,
                               _ptr(ok), _ptr(hr))
    return dict(ttft_ok=bool(ok[0]), tpot_ok=bool(ok[1]), valid=bool(ok[2]), headroom=float(hr[0]),
                ttft_headroom=float(hr[1]))


def score_latency_info(lp: LatencyParams, snap: SnapshotData, have_info, ttft_headroom, tpot_headroom, dispatched=None,
                       mask=None, match=None, total=0):
    out = np.full(snap.M, np.nan)
    hv = _arr(have_info, np.uint8)
    th, ph = _arr(ttft_headroom, np.float64), _arr(tpot_headroom, np.float64)
    dp = _arr(dispatched, np.int32)
    m, mt = _arr(mask, np.uint32), _arr(match, np.uint16)
    lib().orc_score_latency_info(C.byref(lp), C.byref(snap.struct), _ptr(m), _ptr(hv), _ptr(th), _ptr(ph), _ptr(dp),
                                 _ptr(mt), total, _ptr(out))
    return out


def score_latency(lp: LatencyParams, snap: SnapshotData, input_tokens=0, ttft_slo=0.0, tpot_slo=0.0, mask=None,
                  match=None, total=0):
    out = np.full(snap.M, np.nan)
    pred = np.zeros((max(snap.M, 1), 2))
    lr = LatencyRequest(input_tokens, ttft_slo, tpot_slo)
    m, mt = _arr(mask, np.uint32), _arr(match, np.uint16)
    lib().orc_score_latency(C.byref(lp), C.byref(snap.struct), _ptr(m), _ptr(mt), total, C.byref(lr), _ptr(out),
                            _ptr(pred))
    return out, pred[:snap.M]


def score_single(kind: str, snap: SnapshotData, mask=None, adapter_id: int = -1, threshold: float = 0.0):
    out = np.full(snap.M, np.nan)
    m = _arr(mask, np.uint32)
    if kind == "token_load":
        lib().orc_score_token_load(C.byref(snap.struct), _ptr(m), threshold, _ptr(out))
    elif kind == "lora":
        lib().orc_score_lora(C.byref(snap.struct), _ptr(m), adapter_id, _ptr(out))
    else:
        getattr(lib(), f"orc_score_{kind}")(C.byref(snap.struct), _ptr(m), _ptr(out))
    return out


def score_prefix(match, total, M, have_info=True):
    out = np.full(M, np.nan)
    m = _arr(match, np.uint16)
    lib().orc_score_prefix(M, None, _ptr(m), total, 1 if have_info else 0, _ptr(out))
    return out


def schedule_one(snap: SnapshotData, profile: Profile, request_index=0, adapter_id=-1, mask=None, match=None,
                 total=0, pair_col=None):
    M = snap.M
    mw = (M + 31) // 32
    pick, score, ties = C.c_int32(), C.c_double(), C.c_int32()
    tie_set = np.zeros(max(mw, 1), dtype=np.uint32)
    weighted = np.zeros(max(M, 1), dtype=np.float64)
    m = _arr(mask, np.uint32)
    mt = _arr(match, np.uint16)
    pc = _arr(pair_col, np.float32)
    rc = lib().orc_schedule_one(C.byref(snap.struct), C.byref(profile), request_index, adapter_id, _ptr(m), _ptr(mt),
                                total, _ptr(pc), C.byref(pick), C.byref(score), C.byref(ties), _ptr(tie_set),
                                _ptr(weighted))
    ts = [i for i in range(M) if (tie_set[i >> 5] >> (i & 31)) & 1]
    return dict(rc=rc, pick=pick.value, score=score.value, tie_count=ties.value, tie_set=ts, weighted=weighted[:M])


def schedule_batch(snap: SnapshotData, profile: Profile, index: Index | None, R: int, *, prompt_bytes=None,
                   prompt_off=None, model_seed=None, hashes_in=None, n_hashes_in=None, adapter_id=None,
                   cand_mask=None, dense_feat=None, dense_total=None, block_chars=64, max_blocks=256,
                   request_base=0, n_threads=1, want_match=False, want_hashes=False, want_tie_set=False,
                   want_scores=False, input_tokens=None, ttft_slo=None, tpot_slo=None, want_pred=False, goshape=False,
                   shuffle_seed=0, want_filter_mask=False):
    M = snap.M
    mw = (M + 31) // 32
    b = Batch()
    keep = []

    def put(name, a, dt):
        a = _arr(a, dt)
        keep.append(a)
        setattr(b, name, _ptr(a))
        return a

    b.R = R
    b.request_base = request_base
    put("prompt_bytes", prompt_bytes, np.uint8)
    put("prompt_off", prompt_off, np.int64)
    put("model_seed", model_seed, np.uint64)
    hi = put("hashes_in", hashes_in, np.uint64)
    put("n_hashes_in", n_hashes_in, np.uint16)
    b.hash_stride = hi.shape[1] if hi is not None else 0
    put("adapter_id", adapter_id, np.int32)
    put("cand_mask", cand_mask, np.uint32)
    put("dense_feat", dense_feat, np.float32)
    put("dense_total", dense_total, np.uint16)
    put("input_tokens", input_tokens, np.int32)
    put("ttft_slo", ttft_slo, np.float64)
    put("tpot_slo", tpot_slo, np.float64)
    b.block_chars = block_chars
    b.max_blocks = max_blocks
    out = dict(pick=np.zeros(R, np.int32), pick_score=np.zeros(R, np.float64), tie_count=np.zeros(R, np.int32),
               total_blocks=np.zeros(R, np.uint16))
    if want_match:
        out["match_blocks"] = np.zeros((R, M), np.uint16)
    if want_hashes:
        out["hashes_out"] = np.zeros((R, max_blocks), np.uint64)
    if want_tie_set:
        out["tie_set"] = np.zeros((R, mw), np.uint32)
    if want_scores:
        out["weighted_out"] = np.zeros((R, M), np.float64)
    if want_pred:
        out["pred_out"] = np.zeros((R, M, 2), np.float64)
    if want_filter_mask:
        out["filter_mask_out"] = np.zeros((R, max(mw, 1)), np.uint32)
    for k, v in out.items():
        setattr(b, k, _ptr(v))
    if goshape:
        rc = goshape_lib().orc_goshape_schedule_batch(C.byref(snap.struct), C.byref(profile),
                                                      index._h if index is not None else None, C.byref(b), n_threads, shuffle_seed)
        assert rc == 0, "goshape baseline covers the unfiltered prompt path only"
        return out
    lib().orc_schedule_batch(C.byref(snap.struct), C.byref(profile), index._h if index is not None else None,
                             C.byref(b), n_threads)
    return out
