"""Thin Python face of the C ABI (include/eppscore.h).  No arithmetic happens here: numpy arrays /
torch tensors are only handed to libeppscore.so as pointers."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi
from ._capi import Batch, Config, EppscoreError, LatencyParams, SCORER, Snapshot, Stats


def default_config(scorers=None, filters=(), **kw) -> Config:
    """scorers: list of (kind name | int, weight) in profile order; default = reference default config
    (queue 2, kv 2, prefix 3; pkg/epp/config/loader/defaults.go:46-103)."""
    cfg = Config()
    capi.lib().eppscore_config_default(C.byref(cfg))
    if scorers is not None:
        if len(scorers) > capi.MAX_SCORERS:
            raise ValueError("too many scorers")
        cfg.n_scorers = len(scorers)
        for i, (k, w) in enumerate(scorers):
            cfg.scorer_kind[i] = SCORER[k] if isinstance(k, str) else int(k)
            cfg.scorer_weight[i] = float(w)
    cfg.n_filters = len(filters)  # [(kind, (params...))] device-side filters, in order
    for i, (k, par) in enumerate(filters):
        cfg.filter_kind[i] = int(k)
        for j, v in enumerate(par):
            cfg.filter_param[i][j] = float(v)
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise TypeError(f"unknown config field {k}")
        setattr(cfg, k, v)
    return cfg


def latency_params(**kw) -> LatencyParams:
    """eppscore_latency_params with the reference defaults (predictedlatency/plugin.go:128-136,
    scorer/latency/plugin.go:83-90); keyword arguments override fields."""
    p = LatencyParams()
    capi.lib().eppscore_latency_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown latency field {k}")
        setattr(p, k, v)
    return p


def _is_torch(x) -> bool:
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda")


class _Args:
    """Keeps converted arrays alive for the duration of a call and yields raw pointers."""

    def __init__(self, device: bool):
        self.device = device
        self.keep = []

    def ptr(self, x, dtype):
        if x is None:
            return None
        if self.device:
            if isinstance(x, int):
                return x
            if not _is_torch(x) or not x.is_cuda:
                raise TypeError("device-location call needs CUDA tensors or raw device pointers")
            if not x.is_contiguous():
                raise ValueError("tensor must be contiguous")
            self.keep.append(x)
            return x.data_ptr()
        if _is_torch(x):
            x = x.numpy()
        a = np.ascontiguousarray(x, dtype=dtype)
        self.keep.append(a)
        return a.ctypes.data


class Engine:
    """One engine per CUDA device (eppscore_create)."""

    def __init__(self, config: Config | None = None, device: int = 0):
        self._lib = capi.lib()
        self.cfg = config if config is not None else default_config()
        h = C.c_void_p()
        rc = self._lib.eppscore_create(device, C.byref(self.cfg), C.byref(h))
        if rc != 0:
            raise EppscoreError(rc, (self._lib.eppscore_last_error(None) or b"").decode())
        self._h = h
        self.M = 0

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_h", None):
            self._lib.eppscore_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise EppscoreError(rc, (self._lib.eppscore_last_error(self._h) or b"").decode())
        return rc

    def stats(self) -> Stats:
        s = Stats()
        self._check(self._lib.eppscore_get_stats(self._h, C.byref(s)))
        return s

    def set_debug(self, key: int, value: int):
        self._check(self._lib.eppscore_set_debug(self._h, key, value))

    def set_latency_params(self, params: LatencyParams):
        """Takes effect at the next set_snapshot."""
        self._check(self._lib.eppscore_set_latency_params(self._h, C.byref(params)))

    # ------------------------------------------------------------------ snapshot
    def set_snapshot(self, kv_usage, queue, running=None, lora_active=None, lora_waiting=None, lora_nmodels=None,
                     lora_max=None, endpoint_cols=(), epoch=0, device=False, stream=None, M=None, lora_words=None,
                     min_tpot_slo=None, dispatched=None, prefill_role=None, inflight_tokens=None):
        a = _Args(device)
        s = Snapshot()
        s.struct_size = C.sizeof(Snapshot)
        if M is None:
            M = int(kv_usage.shape[0]) if hasattr(kv_usage, "shape") else len(kv_usage)
        s.M = M
        if lora_words is None:
            if lora_active is None:
                lora_words = 0
            else:
                n = int(np.prod(lora_active.shape)) if hasattr(lora_active, "shape") else len(lora_active)
                lora_words = (n // M) if M else 1
        s.lora_words = lora_words
        s.location = 1 if device else 0
        s.kv_usage = a.ptr(kv_usage, np.float64)
        s.queue = a.ptr(queue, np.int64)
        s.running = a.ptr(running, np.int64)
        s.lora_active = a.ptr(lora_active, np.uint64)
        s.lora_waiting = a.ptr(lora_waiting, np.uint64)
        s.lora_nmodels = a.ptr(lora_nmodels, np.int32)
        s.lora_max = a.ptr(lora_max, np.int32)
        for i, c in enumerate(endpoint_cols):
            s.endpoint_col[i] = a.ptr(c, np.float64)
        s.epoch = epoch
        s.stream = stream
        s.min_tpot_slo = a.ptr(min_tpot_slo, np.float64)
        s.dispatched = a.ptr(dispatched, np.int32)
        s.prefill_role = a.ptr(prefill_role, np.uint8)
        s.inflight_tokens = a.ptr(inflight_tokens, np.int64)
        self._check(self._lib.eppscore_set_snapshot(self._h, C.byref(s)))
        self.M = M

    # ------------------------------------------------------------------ the hot path
    def schedule(self, R, *, prompt_bytes=None, prompt_off=None, prompt_len=None, model_seed=None, hashes_in=None,
                 n_hashes_in=None, hash_stride=0, adapter_id=None, cand_mask=None, dense_feat=None, dense_total=None,
                 block_chars=0, max_blocks=0, request_base=0, want_match=False, want_total=True, want_hashes=False,
                 want_scores=False, input_tokens=None, ttft_slo=None, tpot_slo=None, want_pred=False, want_filter_mask=False,
                 device=False, stream=None, out=None):
        """Host mode (device=False): numpy in, returns dict of numpy outputs (copies inside the call).
        Device mode: CUDA tensors / raw pointers in; `out` must hold preallocated CUDA tensors
        (pick int32[R], pick_score float64[R], tie_count int32[R], optional match_blocks/total_blocks/hashes_out);
        the call is asynchronous on `stream`."""
        a = _Args(device)
        b = Batch()
        b.struct_size = C.sizeof(Batch)
        b.R = R
        b.location = 1 if device else 0
        b.request_base = request_base
        b.prompt_bytes = a.ptr(prompt_bytes, np.uint8)
        b.prompt_off = a.ptr(prompt_off, np.int64)
        b.prompt_len = a.ptr(prompt_len, np.int32)
        b.model_seed = a.ptr(model_seed, np.uint64)
        b.hashes_in = a.ptr(hashes_in, np.uint64)
        b.n_hashes_in = a.ptr(n_hashes_in, np.uint16)
        if hashes_in is not None and not hash_stride:
            hash_stride = int(hashes_in.shape[1])
        b.hash_stride = hash_stride
        b.block_chars = block_chars
        b.max_blocks = max_blocks
        b.adapter_id = a.ptr(adapter_id, np.int32)
        b.cand_mask = a.ptr(cand_mask, np.uint32)
        b.dense_feat = a.ptr(dense_feat, np.float32)
        b.dense_total = a.ptr(dense_total, np.uint16)
        b.input_tokens = a.ptr(input_tokens, np.int32)
        b.ttft_slo = a.ptr(ttft_slo, np.float64)
        b.tpot_slo = a.ptr(tpot_slo, np.float64)
        b.stream = stream
        if device:
            if out is None:
                raise ValueError("device mode needs preallocated outputs")
            b.pick = a.ptr(out["pick"], None)
            b.pick_score = a.ptr(out["pick_score"], None)
            b.tie_count = a.ptr(out["tie_count"], None)
            b.match_blocks = a.ptr(out.get("match_blocks"), None)
            b.total_blocks = a.ptr(out.get("total_blocks"), None)
            b.hashes_out = a.ptr(out.get("hashes_out"), None)
            b.scores_out = a.ptr(out.get("scores_out"), None)
            b.pred_out = a.ptr(out.get("pred_out"), None)
            b.filter_mask_out = a.ptr(out.get("filter_mask_out"), None)
            self._check(self._lib.eppscore_schedule_batch(self._h, C.byref(b)))
            return out
        M = self.M
        mb = max_blocks or self.cfg.max_blocks
        res = dict(pick=np.full(R, -2, np.int32), pick_score=np.zeros(R, np.float64), tie_count=np.zeros(R, np.int32))
        if want_total:
            res["total_blocks"] = np.zeros(R, np.uint16)
        if want_match:
            res["match_blocks"] = np.zeros((R, M), np.uint16)
        if want_hashes:
            res["hashes_out"] = np.zeros((R, mb), np.uint64)
        if want_scores:
            res["scores_out"] = np.zeros((R, M), np.float64)
        if want_pred:
            res["pred_out"] = np.zeros((R, M, 2), np.float64)
        if want_filter_mask:
            res["filter_mask_out"] = np.zeros((R, max((M + 31) // 32, 1)), np.uint32)
        for k, v in res.items():
            setattr(b, k, v.ctypes.data)
        self._check(self._lib.eppscore_schedule_batch(self._h, C.byref(b)))
        return res

    def hash_prompts(self, prompt_bytes, prompt_off, model_seed=None, prompt_len=None, block_chars=0, max_blocks=0):
        R = len(prompt_off) - 1
        mb = max_blocks or self.cfg.max_blocks
        a = _Args(False)
        hashes = np.zeros((max(R, 1), mb), np.uint64)
        n = np.zeros(max(R, 1), np.uint16)
        self._check(self._lib.eppscore_hash_prompts(
            self._h, R, 0, a.ptr(prompt_bytes, np.uint8), a.ptr(prompt_off, np.int64), a.ptr(prompt_len, np.int32),
            a.ptr(model_seed, np.uint64), block_chars, max_blocks, hashes.ctypes.data, n.ctypes.data, None))
        return hashes[:R], n[:R]

    @staticmethod
    def hash_prompts_host(prompt_bytes, prompt_off, model_seed=None, prompt_len=None, block_chars=64, max_blocks=256, stride=0,
                          n_threads=0, out=None):
        """hashPrompt on the host cores (library worker pool).  Returns (hashes [R, stride], n_hashes [R])."""
        R = len(prompt_off) - 1
        stride = stride or max_blocks
        a = _Args(False)
        if out is None:
            out = (np.empty((max(R, 1), stride), np.uint64), np.empty(max(R, 1), np.uint16))
        rc = capi.lib().eppscore_hash_prompts_host(R, a.ptr(prompt_bytes, np.uint8), a.ptr(prompt_off, np.int64),
                                                   a.ptr(prompt_len, np.int32), a.ptr(model_seed, np.uint64), block_chars, max_blocks,
                                                   out[0].ctypes.data, stride, out[1].ctypes.data, n_threads)
        if rc != 0:
            raise EppscoreError(rc, "eppscore_hash_prompts_host")
        return out[0][:R], out[1][:R]

    def count_fields(self, prompt_bytes, prompt_off, prompt_len=None):
        """len(strings.Fields(prompt)) per request, computed on the device (host arrays in and out)."""
        R = len(prompt_off) - 1
        a = _Args(False)
        out = np.zeros(max(R, 1), np.int32)
        self._check(self._lib.eppscore_count_fields(self._h, R, 0, a.ptr(prompt_bytes, np.uint8), a.ptr(prompt_off, np.int64),
                                                    a.ptr(prompt_len, np.int32), out.ctypes.data, None))
        return out[:R]

    @staticmethod
    def model_seed(model, salt=b"") -> int:
        if isinstance(model, str):
            model = model.encode()
        if isinstance(salt, str):
            salt = salt.encode()
        return capi.lib().eppscore_model_seed(model, len(model), salt, len(salt))

    # ------------------------------------------------------------------ prefix index
    def commit_picks(self, pick, hashes, n_hashes, lru_capacity=None):
        a = _Args(False)
        hashes = np.ascontiguousarray(hashes, np.uint64)
        stride = hashes.shape[1] if hashes.ndim == 2 else 0
        self._check(self._lib.eppscore_commit_picks(self._h, len(pick), a.ptr(pick, np.int32), a.ptr(hashes, np.uint64),
                                                    a.ptr(n_hashes, np.uint16), stride, a.ptr(lru_capacity, np.int32)))

    def commit_picks_device(self, pick, hashes, n_hashes, lru_capacity=None, touch_bound=0, stream=None):
        """PreRequest with DEVICE arrays (CUDA tensors / raw pointers), asynchronous and ordered on `stream`:
        pick int32[R], hashes uint64[R, stride], n_hashes uint16[R]; lru_capacity is a HOST array."""
        a = _Args(True)
        ah = _Args(False)
        R = int(pick.shape[0])
        stride = int(hashes.shape[1])
        self._check(self._lib.eppscore_commit_picks_device(self._h, R, a.ptr(pick, None), a.ptr(hashes, None), a.ptr(n_hashes, None),
                                                           stride, ah.ptr(lru_capacity, np.int32), int(touch_bound), stream))

    def prefix_add(self, hashes, endpoint, lru_capacity=0):
        a = _Args(False)
        h = np.ascontiguousarray(hashes, np.uint64)
        self._check(self._lib.eppscore_prefix_add(self._h, a.ptr(h, np.uint64), len(h), endpoint, lru_capacity))

    def prefix_apply(self, hash_, endpoint, op):
        a = _Args(False)
        self._check(self._lib.eppscore_prefix_apply(self._h, len(hash_), a.ptr(hash_, np.uint64),
                                                    a.ptr(endpoint, np.int32), a.ptr(op, np.uint8)))

    def prefix_remove_endpoint(self, endpoint):
        self._check(self._lib.eppscore_prefix_remove_endpoint(self._h, endpoint))

    def prefix_get(self, hash_: int) -> set:
        words = (self.cfg.max_endpoints + 31) // 32 + 1
        words = max(words, 8)
        # geometry may pad beyond max_endpoints; ask for a generous number of words
        buf = np.zeros(512, np.uint32)
        n = self._check(self._lib.eppscore_prefix_get(self._h, int(hash_), buf.ctypes.data, len(buf)))
        s = {w * 32 + b for w in range(len(buf)) if buf[w] for b in range(32) if (int(buf[w]) >> b) & 1}
        assert len(s) == n, (s, n)
        return s

    def prefix_lru_len(self, endpoint) -> int:
        return self._lib.eppscore_prefix_lru_len(self._h, endpoint)

    def prefix_lru_keys(self, endpoint):
        n = self.prefix_lru_len(endpoint)
        if n < 0:
            return None
        out = np.zeros(max(n, 1), np.uint64)
        self._lib.eppscore_prefix_lru_keys(self._h, endpoint, out.ctypes.data, len(out))
        return [int(x) for x in out[:n]]
