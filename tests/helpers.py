"""Shared helpers for the parity tests: packing the reference's map-shaped LoRA state into the
mask-shaped snapshot both the oracle and the engine consume, and seeded synthetic workloads
(SURVEY.md §8d)."""
from __future__ import annotations

import numpy as np

KIND = {"queue": 0, "kv": 1, "prefix": 2, "lora": 3, "running": 4, "col0": 8, "col1": 9, "col2": 10, "col3": 11,
        "pair0": 16, "pair1": 17}


def kinds(scorers):
    return [(KIND[k], float(w)) for k, w in scorers]


def pack_lora(endpoints, target_models=()):
    """endpoints: list of dicts with 'active', 'waiting' (lists of names), 'max_active'.
    Returns (dictionary name->id, active mask [M,W] u64, waiting mask, nmodels i32, max i32).
    nmodels = len(ActiveModels)+len(WaitingModels) as MAP sizes (lora_affinity.go:90): an adapter
    present in both maps counts twice."""
    names = []
    for e in endpoints:
        for n in list(e.get("active", [])) + list(e.get("waiting", [])):
            if n not in names:
                names.append(n)
    for n in target_models:
        if n not in names:
            names.append(n)
    ids = {n: i for i, n in enumerate(names)}
    W = max(1, (len(names) + 63) // 64)
    M = len(endpoints)
    act = np.zeros((M, W), dtype=np.uint64)
    wai = np.zeros((M, W), dtype=np.uint64)
    nm = np.zeros(M, dtype=np.int32)
    mx = np.zeros(M, dtype=np.int32)
    for m, e in enumerate(endpoints):
        for n in set(e.get("active", [])):
            act[m, ids[n] >> 6] |= np.uint64(1) << np.uint64(ids[n] & 63)
        for n in set(e.get("waiting", [])):
            wai[m, ids[n] >> 6] |= np.uint64(1) << np.uint64(ids[n] & 63)
        nm[m] = len(set(e.get("active", []))) + len(set(e.get("waiting", [])))
        mx[m] = e.get("max_active", 0)
    return ids, act, wai, nm, mx


def mask_from_list(M, keep):
    mw = (M + 31) // 32
    mask = np.zeros(max(mw, 1), dtype=np.uint32)
    for m in keep:
        mask[m >> 5] |= np.uint32(1 << (m & 31))
    return mask


def synth_snapshot(M, A=64, seed=0, tie_heavy=False):
    """SURVEY §8d synthetic snapshot."""
    rng = np.random.Generator(np.random.PCG64(seed))
    kv = rng.random(M)
    if tie_heavy:
        kv = np.round(kv, 2)
    queue = np.clip(rng.poisson(8, M), 0, 255).astype(np.int64)
    running = rng.poisson(16, M).astype(np.int64)
    W = max(1, (A + 63) // 64)
    act = np.zeros((M, W), dtype=np.uint64)
    wai = np.zeros((M, W), dtype=np.uint64)
    nm = np.zeros(M, dtype=np.int32)
    for m in range(M):
        ka = int(rng.integers(0, 9))
        kw = int(rng.integers(0, 3))
        ids = rng.choice(A, size=min(A, ka + kw), replace=False)
        for a in ids[:ka]:
            act[m, a >> 6] |= np.uint64(1) << np.uint64(a & 63)
        for a in ids[ka:]:
            wai[m, a >> 6] |= np.uint64(1) << np.uint64(a & 63)
        nm[m] = len(ids) + int(rng.integers(0, 2))  # + out-of-vocabulary adapter
    mx = np.full(M, 8, dtype=np.int32)
    return dict(kv_usage=kv, queue=queue, running=running, lora_active=act, lora_waiting=wai, lora_nmodels=nm,
                lora_max=mx)


def synth_prompts(R, prompt_len=2048, groups=150, shared=1024, seed=0, prefix_seed=None):
    """Shared-prefix prompts: one of `groups` x `shared`-byte prefixes + unique tail, bytes a..z.
    The prefix pool depends only on prefix_seed (default: seed), so batches with different `seed`
    but the same `prefix_seed` share their system prompts — what makes the prefix index hit."""
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    shared = min(shared, prompt_len)
    prng = np.random.Generator(np.random.PCG64((seed if prefix_seed is None else prefix_seed) + 5000))
    prefixes = prng.integers(97, 123, size=(groups, shared), dtype=np.uint8)
    g = rng.integers(0, groups, size=R)
    out = np.empty((R, prompt_len), dtype=np.uint8)
    out[:, :shared] = prefixes[g]
    if prompt_len > shared:
        out[:, shared:] = rng.integers(97, 123, size=(R, prompt_len - shared), dtype=np.uint8)
    off = np.arange(R + 1, dtype=np.int64) * prompt_len
    return out.reshape(-1), off, g


def synth_ragged_prompts(R, max_len=700, seed=0, groups=20, shared=256):
    rng = np.random.Generator(np.random.PCG64(seed + 2000))
    prefixes = rng.integers(97, 123, size=(groups, shared), dtype=np.uint8)
    lens = rng.integers(0, max_len + 1, size=R)
    lens[: min(R, 4)] = [0, 1, 63, 64][: min(R, 4)]
    chunks = []
    for r in range(R):
        n = int(lens[r])
        p = np.concatenate([prefixes[rng.integers(0, groups)], rng.integers(97, 123, size=max_len, dtype=np.uint8)])[:n]
        chunks.append(p)
    off = np.zeros(R + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    data = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint8)
    if data.size == 0:
        data = np.zeros(1, dtype=np.uint8)
    return data.astype(np.uint8), off


def zipf_adapters(R, A=64, seed=0, none_frac=0.05):
    rng = np.random.Generator(np.random.PCG64(seed + 3000))
    p = 1.0 / np.arange(1, A + 1) ** 1.1
    p /= p.sum()
    a = rng.choice(A, size=R, p=p).astype(np.int32)
    a[rng.random(R) < none_frac] = -1
    return a
