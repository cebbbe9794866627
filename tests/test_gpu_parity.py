"""GPU parity tests: every result of the CUDA engine, obtained THROUGH THE C ABI (libeppscore.so via
ctypes), is compared with the CPU oracle on the same seeded inputs — bit-exact for hashes, match
counts, picks, tie counts and float64 scores — and with the reference's own golden vectors."""
import numpy as np
import pytest

import _pkg
from oracle import oracle_py as o
from tests.helpers import (kinds, mask_from_list, pack_lora, synth_prompts, synth_ragged_prompts, synth_snapshot,
                           zipf_adapters)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    _pkg.load_build().build()
    return _pkg.load()


def make_engine(pkg, scorers, M, **kw):
    kw.setdefault("prefix_capacity", 1 << 14)
    return pkg.Engine(pkg.default_config(scorers, max_endpoints=max(M, 1), **kw))


def profile_of(pkg, scorers, tie_mode=0, tie_seed=0):
    return o.make_profile([(pkg.SCORER[k], w) for k, w in scorers], tie_mode=tie_mode, tie_seed=tie_seed)


def assert_same(got, want, keys=("pick", "pick_score", "tie_count")):
    for k in keys:
        g, w = got[k], want["weighted_out" if k == "scores_out" else k]
        if g.dtype.kind == "f":
            # bit-exact except that NaN payloads (non-candidates) may differ
            both = ~(np.isnan(g) | np.isnan(w))
            assert np.array_equal(np.isnan(g), np.isnan(w)), k
            assert np.array_equal(g[both].view(np.uint64), w[both].view(np.uint64)), k
        else:
            assert np.array_equal(g, w), k


# ------------------------------------------------------------------------------------------ hashing
def test_hash_chain_golden_vectors(pkg, xxh_kat):
    """hashPrompt on the GPU against the XXH64 known-answer chains (block sizes 4..256, partial blocks,
    truncation), both with unaligned back-to-back prompts and with 16-byte aligned starts."""
    eng = make_engine(pkg, [("prefix", 1.0)], 8)
    by_cfg = {}
    for c in xxh_kat["chains"]:
        by_cfg.setdefault((c["block_chars"], c["max_blocks"]), []).append(c)
    for (bc, mb), cases in by_cfg.items():
        prompts = [bytes.fromhex(c["prompt_hex"]) for c in cases]
        seeds = np.array([int(c["seed"], 16) for c in cases], np.uint64)
        for aligned in (False, True):
            off = np.zeros(len(prompts) + 1, np.int64)
            lens = np.array([len(p) for p in prompts], np.int32)
            buf = bytearray()
            for i, p in enumerate(prompts):
                if aligned:
                    buf.extend(b"\0" * ((-len(buf)) % 16))
                off[i] = len(buf)
                buf.extend(p)
            off[-1] = len(buf)
            data = np.frombuffer(bytes(buf) + b"\0", np.uint8)
            hashes, n = eng.hash_prompts(data, off, seeds, prompt_len=lens, block_chars=bc, max_blocks=mb)
            for i, c in enumerate(cases):
                assert [f"{int(x):016x}" for x in hashes[i, : n[i]]] == c["hashes"], (bc, mb, aligned, i)
    eng.close()


def test_hash_matches_oracle_ragged(pkg):
    eng = make_engine(pkg, [("prefix", 1.0)], 8)
    for bc, mb, seed in ((64, 256, 1), (64, 4, 2), (32, 256, 3), (4, 256, 4), (100, 7, 5), (128, 2, 6)):
        R = 257
        data, off = synth_ragged_prompts(R, max_len=900, seed=seed)
        seeds = np.arange(R, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)
        hashes, n = eng.hash_prompts(data, off, seeds, block_chars=bc, max_blocks=mb)
        for r in range(R):
            want = o.hash_prompt(bytes(data[off[r]:off[r + 1]]), int(seeds[r]), bc, mb)
            assert n[r] == len(want), (bc, mb, r)
            assert np.array_equal(hashes[r, : n[r]], want), (bc, mb, r)
    eng.close()


# ------------------------------------------------------------------------------------------ golden vectors
def _snapshot_from_endpoints(eps, target_models=()):
    ids, act, wai, nm, mx = pack_lora(eps, target_models)
    sd = dict(kv_usage=np.array([e.get("kv", 0.0) for e in eps], np.float64),
              queue=np.array([e.get("queue", 0) for e in eps], np.int64),
              lora_active=act, lora_waiting=wai, lora_nmodels=nm, lora_max=mx)
    return sd, ids


def test_schedule_finds_optimal_endpoint(pkg, golden):
    """pkg/epp/scheduling/scheduler_test.go:86-143 — pod2 with Score == 2.8 exactly."""
    c = golden["schedule"][0]
    sd, ids = _snapshot_from_endpoints(c["endpoints"], [c["target_model"]])
    eng = make_engine(pkg, c["scorers"], len(c["endpoints"]))
    eng.set_snapshot(**sd)
    res = eng.schedule(1, adapter_id=np.array([ids[c["target_model"]]], np.int32), want_scores=True)
    assert c["endpoints"][res["pick"][0]]["name"] == c["want_pick"]
    assert res["pick_score"][0] == c["want_score_exact"]
    assert res["tie_count"][0] == 1
    # the whole weighted map, as the Go test would see it in DEBUG logs: 0.8+1+0+0 , 0.8+1+0+1, 0.2+0+0+0.8
    assert list(res["scores_out"][0]) == [1.8, 2.8, 1.0]
    eng.close()


def test_no_candidates_is_an_error_pick(pkg, golden):
    """scheduler_test.go:69-78 / scheduler_profile.go:119-121: empty candidate set ⇒ no pick."""
    eng = make_engine(pkg, [("kv", 1), ("queue", 1)], 4)
    eng.set_snapshot(np.array([0.1, 0.2, 0.3, 0.4]), np.array([1, 2, 3, 4]))
    res = eng.schedule(2, cand_mask=np.array([[0], [0b0100]], np.uint32))
    assert res["pick"][0] == -1 and res["tie_count"][0] == 0
    assert res["pick"][1] == 2 and res["tie_count"][1] == 1
    eng.close()


def test_weighted_constant_scorers_and_clamp(pkg, golden):
    """scheduler_profile_test.go:63-108,350-413 — weight math exact (1.1, 50), clamp of out-of-range scores."""
    for c in golden["weighted_constant_scorers"]:
        n = c["n_endpoints"]
        eng = make_engine(pkg, [("col0", c["weights"][0]), ("col1", c["weights"][1])], n)
        eng.set_snapshot(np.zeros(n), np.zeros(n, np.int64), endpoint_cols=[np.full(n, c["scores"][0]), np.full(n, c["scores"][1])])
        res = eng.schedule(1, cand_mask=mask_from_list(n, c["filter_keep"]).reshape(1, -1))
        if c.get("want_error"):
            assert res["pick"][0] == -1 and res["tie_count"][0] == 0
        else:
            assert res["pick_score"][0] == c["want_score_exact"]
            assert res["tie_count"][0] == c["want_tie_count"] and res["pick"][0] in c["filter_keep"]
        eng.close()
    r = golden["enforce_score_range"]["run_out_of_range"]
    eng = make_engine(pkg, [("col0", r["weights"][0]), ("col1", r["weights"][1])], 1)
    eng.set_snapshot(np.zeros(1), np.zeros(1, np.int64), endpoint_cols=[np.array([r["scores"][0]]), np.array([r["scores"][1]])])
    assert eng.schedule(1)["pick_score"][0] == r["want_score_exact"]
    eng.close()
    for x, want in golden["enforce_score_range"]["cases"]:
        eng = make_engine(pkg, [("col0", 1.0)], 1)
        eng.set_snapshot(np.zeros(1), np.zeros(1, np.int64), endpoint_cols=[np.array([x])])
        assert eng.schedule(1)["pick_score"][0] == want
        eng.close()


def test_single_scorer_tables(pkg, golden):
    """The per-scorer table tests of the reference, read back through scores_out (weight 1)."""
    for c in golden["kv_scorer"]["cases"]:
        eng = make_engine(pkg, [("kv", 1.0)], len(c["kv"]))
        eng.set_snapshot(np.array(c["kv"]), np.zeros(len(c["kv"]), np.int64))
        got = eng.schedule(1, want_scores=True)["scores_out"][0]
        assert np.allclose(got, c["want"], atol=golden["kv_scorer"]["tolerance"], rtol=0)
        eng.close()
    for c in golden["queue_scorer"]["cases"]:
        eng = make_engine(pkg, [("queue", 1.0)], len(c["queue"]))
        eng.set_snapshot(np.zeros(len(c["queue"])), np.array(c["queue"], np.int64))
        got = eng.schedule(1, want_scores=True)["scores_out"][0]
        assert np.allclose(got, c["want"], atol=golden["queue_scorer"]["tolerance"], rtol=0)
        eng.close()
    for c in golden["running_scorer"]["cases"]:
        n = len(c["running"])
        eng = make_engine(pkg, [("running", 1.0)], n)
        eng.set_snapshot(np.zeros(n), np.zeros(n, np.int64), np.array(c["running"], np.int64))
        got = eng.schedule(1, want_scores=True)["scores_out"][0]
        assert np.allclose(got, c["want"], atol=golden["running_scorer"]["tolerance"], rtol=0)
        eng.close()
    for c in golden["lora_scorer"]["cases"]:
        if not c["endpoints"]:
            continue
        sd, ids = _snapshot_from_endpoints(c["endpoints"], [c["target"]])
        eng = make_engine(pkg, [("lora", 1.0)], len(c["endpoints"]))
        eng.set_snapshot(**sd)
        got = eng.schedule(1, adapter_id=np.array([ids[c["target"]]], np.int32), want_scores=True)["scores_out"][0]
        assert np.allclose(got, c["want"], atol=golden["lora_scorer"]["tolerance"], rtol=0), c["name"]
        eng.close()
    for c in golden["prefix_scorer"]["cases"]:
        n = len(c["match"])
        eng = make_engine(pkg, [("prefix", 1.0)], n)
        eng.set_snapshot(np.zeros(n), np.zeros(n, np.int64))
        feat = np.zeros((1, n, 4), np.float32)
        feat[0, :, 0] = c["match"]
        got = eng.schedule(1, dense_feat=feat, dense_total=np.array([c["total"]], np.uint16), want_scores=True)
        assert list(got["scores_out"][0]) == c["want_exact"]  # assert.Equal in the reference
        eng.close()


def test_integration_routing(pkg, golden):
    """test/integration/epp routing scenarios (queue+kv+prefix+lora, weight 1, incl. subset masks)."""
    g = golden["integration_routing"]
    for c in g["cases"]:
        scorers = c.get("scorers", g["scorers"])
        eps = [{"queue": q, "kv": kv, "active": models, "waiting": [], "max_active": 0} for _, q, kv, models in c["pods"]]
        sd, ids = _snapshot_from_endpoints(eps, [c["target_model"]])
        eng = make_engine(pkg, scorers, len(eps))
        eng.set_snapshot(**sd)
        prompt = np.frombuffer(c["prompt"].encode(), np.uint8)
        kw = dict(prompt_bytes=prompt, prompt_off=np.array([0, len(prompt)], np.int64),
                  model_seed=np.array([eng.model_seed(c["target_model"])], np.uint64),
                  adapter_id=np.array([ids[c["target_model"]]], np.int32))
        if "subset" in c:
            kw["cand_mask"] = mask_from_list(len(eps), c["subset"]).reshape(1, -1)
        res = eng.schedule(1, **kw)
        if c.get("want_error"):
            assert res["pick"][0] == -1, c["name"]
        else:
            assert res["pick"][0] == c["want_pick"], c["name"]
            assert res["tie_count"][0] == 1, c["name"]
            assert res["total_blocks"][0] == 0  # prompts shorter than one block (hashing.go:57-60)
        eng.close()


def test_picker_vectors(pkg, golden):
    """maxscore/picker_test.go:43-110: arg-max, ties as sets; seeded-random mode stays inside the tie set."""
    for c in golden["picker"]["cases"]:
        n = len(c["scores"])
        col = np.array([s / 100.0 for s in c["scores"]])
        top = max(c["scores"])
        tie_set = [m for m in range(n) if c["scores"][m] == top]
        eng = make_engine(pkg, [("col0", 100.0)], n)
        eng.set_snapshot(np.zeros(n), np.zeros(n, np.int64), endpoint_cols=[col])
        res = eng.schedule(1)
        assert res["pick"][0] == tie_set[0] and res["tie_count"][0] == len(tie_set)
        eng.close()
        eng = make_engine(pkg, [("col0", 100.0)], n, tie_mode=pkg.TIE_SEEDED_RANDOM, tie_seed=7)
        eng.set_snapshot(np.zeros(n), np.zeros(n, np.int64), endpoint_cols=[col])
        res = eng.schedule(64)
        snap = o.SnapshotData(kv_usage=np.zeros(n), queue=np.zeros(n), endpoint_cols=[col])
        want = o.schedule_batch(snap, o.make_profile([(8, 100.0)], tie_mode=1, tie_seed=7), None, 64)
        assert np.array_equal(res["pick"], want["pick"])
        assert set(res["pick"]) == set(tie_set)
        eng.close()


# ------------------------------------------------------------------------------------------ prefix index
def test_prefix_completion_and_prerequest(pkg, golden):
    """approximateprefix/plugin_test.go:37-227: empty index ⇒ 0 of 2; after pick(pod1)+prefill(pod3) of
    "aaaaaa", "aaaabbbb" matches pod1=1, pod3=1, pod2=0 of 2; PreRequest maps every hash to the pick."""
    c = golden["prefix_completion"]
    M = c["n_endpoints"]
    eng = make_engine(pkg, [("prefix", 1.0)], M, block_chars=c["block_chars"], max_blocks=c["max_blocks"])
    eng.set_snapshot(np.zeros(M), np.zeros(M, np.int64))
    seed = np.array([eng.model_seed(c["model"])], np.uint64)

    def run(prompt):
        p = np.frombuffer(prompt.encode(), np.uint8)
        return eng.schedule(1, prompt_bytes=p, prompt_off=np.array([0, len(p)], np.int64), model_seed=seed,
                            want_match=True, want_hashes=True)

    r1 = run(c["first_prompt"])
    assert r1["total_blocks"][0] == c["first_total"] and not r1["match_blocks"].any()
    h1 = r1["hashes_out"][0, : r1["total_blocks"][0]]
    for s in c["commit_to"]:
        eng.prefix_add(h1, s)
    r2 = run(c["second_prompt"])
    assert list(r2["match_blocks"][0]) == c["want_match"] and r2["total_blocks"][0] == c["want_total"]
    for h in h1:
        assert eng.prefix_get(int(h)) == set(c["commit_to"])
    # PreRequest via commit_picks
    eng2 = make_engine(pkg, [("prefix", 1.0)], M, block_chars=4)
    eng2.set_snapshot(np.zeros(M), np.zeros(M, np.int64))
    p = np.frombuffer(golden["pre_request"]["prompt"].encode(), np.uint8)
    r = eng2.schedule(1, prompt_bytes=p, prompt_off=np.array([0, len(p)], np.int64), model_seed=seed, want_hashes=True)
    eng2.commit_picks(np.array([golden["pre_request"]["pick"]], np.int32), r["hashes_out"], r["total_blocks"])
    for h in r["hashes_out"][0, : r["total_blocks"][0]]:
        assert golden["pre_request"]["pick"] in eng2.prefix_get(int(h))
    eng.close()
    eng2.close()


def test_indexer_vectors_on_device(pkg, golden):
    """approximateprefix/indexer_test.go:27-113 with Get() answered by the DEVICE table."""
    c = golden["indexer"]["add_and_get"]
    eng = make_engine(pkg, [("prefix", 1.0)], 4, lru_capacity_default=c["default_lru"])
    for st in c["steps"]:
        eng.prefix_add(st["add"], 0, c["gpu_blocks"])
        assert eng.prefix_lru_len(0) == st["want_len"]
        for h, want in st.get("want_get", {}).items():
            assert sorted(eng.prefix_get(int(h))) == want
    eng.close()
    n = golden["indexer"]["remove_pod_and_eviction"]["indexer_size"]
    eng = make_engine(pkg, [("prefix", 1.0)], 4, lru_capacity_default=n)
    for j in range(n):
        eng.prefix_add([j], 1)
        eng.prefix_add([j], 2)
    for j in range(n):
        assert eng.prefix_get(j) == {1, 2}
    eng.prefix_add([n], 1)
    assert eng.prefix_lru_len(1) == n and eng.prefix_get(0) == {2}
    eng.prefix_remove_endpoint(2)
    assert eng.prefix_get(0) == set() and eng.prefix_lru_len(2) == -1
    for j in range(1, n + 1):
        assert eng.prefix_get(j) == {1}
    assert eng.stats().prefix_live_hashes == n
    assert eng.prefix_lru_keys(1) == list(range(1, n + 1))
    # over-long Add leaves stale hashToPods entries (indexer.go:70-82), like the oracle
    eng2 = make_engine(pkg, [("prefix", 1.0)], 4, lru_capacity_default=2)
    eng2.prefix_add([10, 11, 12], 0)
    assert eng2.prefix_lru_keys(0) == [11, 12] and eng2.prefix_get(10) == {0}
    # raw deltas
    eng2.prefix_apply(np.array([77, 77, 78], np.uint64), np.array([1, 3, 1], np.int32), np.array([0, 0, 0], np.uint8))
    assert eng2.prefix_get(77) == {1, 3}
    eng2.prefix_apply(np.array([77], np.uint64), np.array([1], np.int32), np.array([1], np.uint8))
    assert eng2.prefix_get(77) == {3} and eng2.prefix_get(78) == {1}
    eng.close()
    eng2.close()


def test_commit_replay_matches_oracle_index(pkg):
    """A few rounds of schedule → commit on both sides: LRU state, evictions and match counts stay identical
    (small LRU capacity so evictions, including start-of-chain gaps, really happen)."""
    M, R = 48, 400
    scorers = [("queue", 2), ("kv", 2), ("prefix", 3)]
    eng = make_engine(pkg, scorers, M, lru_capacity_default=150, max_blocks=16, tie_mode=1, tie_seed=3)
    sd = synth_snapshot(M, seed=5, tie_heavy=True)
    eng.set_snapshot(**sd)
    snap = o.SnapshotData(**sd)
    prof = profile_of(pkg, scorers, tie_mode=1, tie_seed=3)
    idx = o.Index(150)
    caps = np.where(np.arange(M) % 3 == 0, 30, 0).astype(np.int32)  # autotune: some endpoints report CacheNumBlocks
    for rnd in range(4):
        prompts, off, _ = synth_prompts(R, prompt_len=640, groups=9, shared=384, seed=20 + rnd % 2, prefix_seed=77)
        seeds = np.full(R, eng.model_seed("m"), np.uint64)
        got = eng.schedule(R, prompt_bytes=prompts, prompt_off=off, model_seed=seeds, want_match=True, want_hashes=True,
                           request_base=rnd * R)
        want = o.schedule_batch(snap, prof, idx, R, prompt_bytes=prompts, prompt_off=off, model_seed=seeds, max_blocks=16,
                                want_match=True, want_hashes=True, request_base=rnd * R)
        assert_same(got, want, ("pick", "pick_score", "tie_count", "total_blocks", "match_blocks"))
        nh = want["total_blocks"]
        for r in range(R):
            assert np.array_equal(got["hashes_out"][r, : nh[r]], want["hashes_out"][r, : nh[r]])
        eng.commit_picks(got["pick"], got["hashes_out"], got["total_blocks"], lru_capacity=caps)
        idx.commit(want["pick"], want["hashes_out"], want["total_blocks"], gpu_blocks=caps)
        for m in range(M):
            assert eng.prefix_lru_keys(m) == idx.lru_keys(m), (rnd, m)
        assert eng.stats().prefix_live_hashes == idx.num_hashes()
        if rnd > 0:
            assert got["match_blocks"].max() > 0
    lens = [eng.prefix_lru_len(m) for m in range(M)]
    assert max(lens) <= 150 and max(lens[0::3]) <= 30  # per-endpoint capacities (default / CacheNumBlocks) respected
    assert eng.stats().lru_entries == sum(x for x in lens if x > 0)
    eng.close()


# ------------------------------------------------------------------------------------------ synthetic parity
CASES = [
    # name, M, R, scorers, kwargs
    ("config_A_cpu_case", 4, 1, [("queue", 1)], {}),
    ("config_B_kv_queue", 256, 4096, [("kv", 1), ("queue", 1)], {}),
    ("config_B_tie_heavy_random", 256, 4096, [("kv", 1), ("queue", 1)], dict(tie_heavy=True, tie_mode=1)),
    ("config_C_prefix", 512, 2048, [("queue", 2), ("kv", 2), ("prefix", 3)], dict(prefix=True)),
    ("config_D_lora_kv", 1024, 4096, [("lora", 1), ("kv", 1)], dict(lora=True)),
    ("headline_all_four", 1024, 4096, [("queue", 2), ("kv", 2), ("prefix", 3), ("lora", 1)], dict(prefix=True, lora=True)),
    ("reference_test_order", 1024, 1024, [("kv", 1), ("queue", 1), ("prefix", 1), ("lora", 1)], dict(prefix=True, lora=True, tie_heavy=True)),
    ("prefix_first_order", 300, 1024, [("prefix", 3), ("queue", 2), ("lora", 1), ("kv", 2)], dict(prefix=True, lora=True)),
    ("odd_M_1000", 1000, 1024, [("queue", 2), ("kv", 2), ("prefix", 3), ("lora", 1)], dict(prefix=True, lora=True)),
    ("M_4096", 4096, 512, [("queue", 2), ("kv", 2), ("prefix", 3), ("lora", 1)], dict(prefix=True, lora=True)),
    ("M_2048_running", 2048, 512, [("running", 1), ("kv", 2), ("prefix", 3)], dict(prefix=True)),
    ("negative_weight", 64, 512, [("queue", -1.5), ("kv", 2), ("lora", 0.25)], dict(lora=True)),
    ("masked_all_four", 1024, 2048, [("queue", 2), ("kv", 2), ("prefix", 3), ("lora", 1)], dict(prefix=True, lora=True, mask=0.5)),
    ("masked_sparse_small", 96, 1024, [("kv", 1), ("queue", 1), ("running", 1)], dict(mask=0.1, tie_heavy=True)),
    ("masked_queue_first_M4096", 4096, 256, [("queue", 1), ("prefix", 1), ("kv", 1)], dict(prefix=True, mask=0.7)),
]


@pytest.mark.parametrize("name,M,R,scorers,opt", CASES, ids=[c[0] for c in CASES])
def test_synthetic_parity(pkg, name, M, R, scorers, opt):
    import zlib
    seed = zlib.crc32(name.encode()) % 1000
    tie_mode = opt.get("tie_mode", 0)
    eng = make_engine(pkg, scorers, M, tie_mode=tie_mode, tie_seed=99, prefix_capacity=1 << 18)
    sd = synth_snapshot(M, seed=seed, tie_heavy=opt.get("tie_heavy", False))
    eng.set_snapshot(**sd)
    snap = o.SnapshotData(**sd)
    prof = profile_of(pkg, scorers, tie_mode=tie_mode, tie_seed=99)
    kw = {}
    idx = None
    if opt.get("prefix"):
        prompts, off, _ = synth_prompts(R, prompt_len=2048, groups=24, shared=1024, seed=seed)
        kw.update(prompt_bytes=prompts, prompt_off=off, model_seed=np.full(R, eng.model_seed("model-x"), np.uint64))
        # warm the index on both sides with an earlier batch routed by the oracle
        idx = o.Index()
        wp, woff, _ = synth_prompts(min(R, 4 * M), prompt_len=2048, groups=24, shared=1024, seed=seed)
        nw = len(woff) - 1
        warm = o.schedule_batch(snap, prof, idx, nw, prompt_bytes=wp, prompt_off=woff,
                                model_seed=np.full(nw, eng.model_seed("model-x"), np.uint64), want_hashes=True, n_threads=8)
        idx.commit(warm["pick"], warm["hashes_out"], warm["total_blocks"])
        eng.commit_picks(warm["pick"], warm["hashes_out"], warm["total_blocks"])
    if opt.get("lora"):
        kw["adapter_id"] = zipf_adapters(R, seed=seed)
    if "mask" in opt:
        rng = np.random.Generator(np.random.PCG64(seed + 7))
        bits = rng.random((R, M)) < opt["mask"]
        bits[0, :] = False  # one request with an empty candidate set
        mask = np.zeros((R, (M + 31) // 32), np.uint32)
        for w in range(mask.shape[1]):
            chunk = bits[:, w * 32:(w + 1) * 32]
            mask[:, w] = (chunk * (1 << np.arange(chunk.shape[1], dtype=np.uint64))).sum(axis=1).astype(np.uint32)
        kw["cand_mask"] = mask
    got = eng.schedule(R, want_match=True, want_scores=True, **kw)
    want = o.schedule_batch(snap, prof, idx, R, want_match=True, want_scores=True, want_tie_set=True, n_threads=8, **kw)
    assert_same(got, want, ("pick", "pick_score", "tie_count", "total_blocks", "match_blocks", "scores_out"))
    # the same call without per-pair diagnostics takes the specialised kernels (pick_sparse / no R x M pass)
    fast = eng.schedule(R, **kw)
    assert_same(fast, want, ("pick", "pick_score", "tie_count", "total_blocks"))
    # reference semantics: the pick is a member of the arg-max set the reference would shuffle over
    ts = want["tie_set"]
    ok = got["pick"] >= 0
    r = np.nonzero(ok)[0]
    assert ((ts[r, got["pick"][r] >> 5] >> (got["pick"][r] & 31).astype(np.uint32)) & 1).all()
    if opt.get("prefix"):
        assert got["match_blocks"].max() > 0
    if "mask" in opt:
        assert got["pick"][0] == -1
    eng.close()


def test_dense_rows_parity(pkg):
    """Dense float4 feature rows {match, lora class, pair0, pair1} streamed from HBM, incl. pair columns."""
    for M, R, masked in ((1024, 2048, False), (333, 1024, True), (4096, 300, False)):
        scorers = [("queue", 2), ("kv", 2), ("prefix", 3), ("lora", 1), ("pair0", 1.5), ("pair1", 0.5)]
        eng = make_engine(pkg, scorers, M, tie_mode=1, tie_seed=5)
        sd = synth_snapshot(M, seed=M)
        eng.set_snapshot(**sd)
        rng = np.random.Generator(np.random.PCG64(M))
        feat = np.zeros((R, M, 4), np.float32)
        total = rng.integers(0, 40, R).astype(np.uint16)
        feat[:, :, 0] = np.minimum(rng.integers(0, 48, (R, M)) * (rng.random((R, M)) < 0.05), total[:, None] + 2)
        feat[:, :, 1] = rng.integers(0, 4, (R, M))
        feat[:, :, 2] = rng.random((R, M)).astype(np.float32) * 1.4 - 0.2  # exercises the clamp
        feat[:, :, 3] = np.round(rng.random((R, M)), 1)
        kw = dict(dense_feat=feat, dense_total=total)
        if masked:
            mask = rng.integers(0, 2 ** 32, (R, (M + 31) // 32), dtype=np.uint64).astype(np.uint32)
            kw["cand_mask"] = mask
        got = eng.schedule(R, want_match=True, want_scores=True, **kw)
        want = o.schedule_batch(o.SnapshotData(**sd), profile_of(pkg, scorers, 1, 5), None, R, want_match=True,
                                want_scores=True, n_threads=8, **kw)
        assert_same(got, want, ("pick", "pick_score", "tie_count", "total_blocks", "match_blocks", "scores_out"))
        fast = eng.schedule(R, **kw)  # template-specialised streaming kernel (unmasked) / generic (masked)
        assert_same(fast, want, ("pick", "pick_score", "tie_count", "total_blocks"))
        eng.close()
    # the specialised sequences of score_dense.cu, each against the oracle
    for scorers in ([("kv", 1)], [("queue", 2), ("kv", 2), ("prefix", 3)], [("kv", 1), ("lora", 1)], [("lora", 1), ("kv", 1)],
                    [("queue", 2), ("kv", 2), ("prefix", 3), ("lora", 1)], [("kv", 1), ("prefix", 3), ("lora", 1), ("pair0", 2)],
                    [("prefix", 1)], [("prefix", 2), ("lora", 1)], [("lora", 1), ("prefix", 2), ("kv", 1)]):
        M, R = 777, 1500
        eng = make_engine(pkg, scorers, M, tie_mode=0)
        sd = synth_snapshot(M, seed=len(scorers), tie_heavy=True)
        eng.set_snapshot(**sd)
        rng = np.random.Generator(np.random.PCG64(len(scorers) + 40))
        feat = np.zeros((R, M, 4), np.float32)
        total = rng.integers(0, 300, R).astype(np.uint16)  # some totals exceed the 256-entry LUT
        feat[:, :, 0] = rng.integers(0, 320, (R, M)) * (rng.random((R, M)) < 0.1)
        feat[:, :, 1] = rng.integers(0, 4, (R, M))
        feat[:, :, 2] = np.round(rng.random((R, M)), 2)
        kw = dict(dense_feat=feat, dense_total=total)
        fast = eng.schedule(R, **kw)
        want = o.schedule_batch(o.SnapshotData(**sd), profile_of(pkg, scorers), None, R, n_threads=8, **kw)
        assert_same(fast, want, ("pick", "pick_score", "tie_count", "total_blocks"))
        eng.close()


def test_sparse_path_stress(pkg):
    """pick_sparse.cu against the oracle where its case analysis is exercised hardest: tie-heavy snapshots,
    hot prefixes cached on MANY endpoints (hundreds of exceptions), both tie modes, several geometries."""
    for M, R, tie_mode, scorers in ((1024, 3000, 0, [("queue", 2), ("kv", 2), ("prefix", 3), ("lora", 1)]),
                                    (1024, 3000, 1, [("queue", 2), ("kv", 2), ("prefix", 3), ("lora", 1)]),
                                    (200, 2000, 1, [("kv", 1), ("prefix", 1)]),
                                    (500, 2000, 0, [("prefix", 0.0), ("kv", 1), ("lora", 2)]),
                                    (3000, 700, 1, [("lora", 1), ("prefix", 0.5), ("kv", 1), ("queue", 1)]),
                                    (6000, 300, 0, [("queue", 2), ("kv", 2), ("prefix", 3)]),
                                    (64, 2000, 1, [("kv", 1), ("queue", 1)]),
                                    (1024, 2000, 1, [("lora", 1), ("kv", 1)])):
        eng = make_engine(pkg, scorers, M, tie_mode=tie_mode, tie_seed=1234, prefix_capacity=1 << 16, max_blocks=64)
        sd = synth_snapshot(M, seed=M + tie_mode, tie_heavy=True)
        sd["kv_usage"] = np.round(sd["kv_usage"], 1)  # very tie heavy
        eng.set_snapshot(**sd)
        snap, prof, idx = o.SnapshotData(**sd), profile_of(pkg, scorers, tie_mode, 1234), o.Index()
        rng = np.random.Generator(np.random.PCG64(M))
        prompts, off, _ = synth_prompts(R, prompt_len=1024, groups=6, shared=512, seed=M, prefix_seed=3)
        seeds = np.full(R, eng.model_seed("s"), np.uint64)
        hashes, nh = eng.hash_prompts(prompts, off, seeds, max_blocks=64)
        # cache group prefixes on many endpoints: group g's first blocks on a random ~30% of the endpoints
        for r in range(0, 60):
            eps = np.nonzero(rng.random(M) < 0.3)[0]
            depth = int(rng.integers(1, 9))
            for ep in eps[:200]:
                eng.prefix_add(hashes[r, :depth], int(ep))
                idx.add(hashes[r, :depth], int(ep))
        ad = zipf_adapters(R, seed=M)
        kw = dict(prompt_bytes=prompts, prompt_off=off, model_seed=seeds, adapter_id=ad, request_base=7000)
        fast = eng.schedule(R, max_blocks=64, **kw)
        want = o.schedule_batch(snap, prof, idx, R, max_blocks=64, want_match=True, n_threads=8, **kw)
        assert_same(fast, want, ("pick", "pick_score", "tie_count", "total_blocks"))
        assert (want["match_blocks"] > 0).sum(axis=1).max() >= min(M, 150) * 0.2  # really many exceptions per request
        eng.close()


def test_sparse_path_full_256_block_match(pkg):
    """defaultMaxPrefixBlocks = 256 matched blocks on an endpoint: the sparse kernel's 8-bit match counters
    wrap to 0 at 256 and must still decode to 256 (score 1.0 * weight)."""
    M, R = 96, 40
    scorers = [("kv", 1), ("prefix", 3)]
    eng = make_engine(pkg, scorers, M, prefix_capacity=1 << 12)
    sd = synth_snapshot(M, seed=12)
    eng.set_snapshot(**sd)
    snap, prof, idx = o.SnapshotData(**sd), profile_of(pkg, scorers), o.Index()
    prompts, off, _ = synth_prompts(R, prompt_len=256 * 64, groups=3, shared=250 * 64, seed=12)
    seeds = np.full(R, eng.model_seed("long"), np.uint64)
    hashes, nh = eng.hash_prompts(prompts, off, seeds)
    assert (nh == 256).all()
    for r, ep in ((0, 5), (1, 70), (2, 5), (3, 33)):   # request r's whole 256-block chain cached on endpoint ep
        eng.prefix_add(hashes[r], ep)
        idx.add(hashes[r], ep)
    kw = dict(prompt_bytes=prompts, prompt_off=off, model_seed=seeds)
    fast = eng.schedule(R, **kw)
    want = o.schedule_batch(snap, prof, idx, R, want_match=True, n_threads=4, **kw)
    assert_same(fast, want, ("pick", "pick_score", "tie_count", "total_blocks"))
    assert want["match_blocks"].max() == 256 and set(fast["pick"][:4]) <= {5, 70, 33}
    full = eng.schedule(R, want_match=True, **kw)
    assert_same(full, want, ("pick", "pick_score", "tie_count", "match_blocks"))
    eng.close()


def test_prefix_table_grows_on_device(pkg):
    """prefix_capacity is only the INITIAL size: the table doubles (host mirror re-hashed, device buffers
    reallocated and re-uploaded) and results stay identical to the oracle."""
    M, R = 200, 1500
    scorers = [("queue", 2), ("kv", 2), ("prefix", 3)]
    eng = make_engine(pkg, scorers, M, prefix_capacity=32)
    sd = synth_snapshot(M, seed=21)
    eng.set_snapshot(**sd)
    snap, prof, idx = o.SnapshotData(**sd), profile_of(pkg, scorers), o.Index()
    prompts, off, _ = synth_prompts(R, prompt_len=1024, groups=12, shared=512, seed=21)
    seeds = np.full(R, eng.model_seed("grow"), np.uint64)
    kw = dict(prompt_bytes=prompts, prompt_off=off, model_seed=seeds)
    for rnd in range(2):
        got = eng.schedule(R, want_hashes=True, **kw)
        want = o.schedule_batch(snap, prof, idx, R, want_hashes=True, n_threads=8, **kw)
        assert_same(got, want, ("pick", "pick_score", "tie_count", "total_blocks"))
        eng.commit_picks(got["pick"], got["hashes_out"], got["total_blocks"])
        idx.commit(want["pick"], want["hashes_out"], want["total_blocks"])
    st = eng.stats()
    assert st.prefix_capacity >= 8192 and st.prefix_live_hashes == idx.num_hashes() > 32
    h0 = int(got["hashes_out"][5, 0])
    assert eng.prefix_get(h0) == idx.get(h0) and len(idx.get(h0)) >= 1
    eng.close()


def test_hashes_in_path_and_small_blocks(pkg):
    """Pre-hashed input (a host that hashes itself) and the generic (block_chars=4, unaligned) hash path."""
    M, R = 128, 600
    scorers = [("prefix", 3), ("kv", 1)]
    eng = make_engine(pkg, scorers, M, block_chars=4, max_blocks=300, prefix_capacity=1 << 17)
    sd = synth_snapshot(M, seed=8)
    eng.set_snapshot(**sd)
    data, off = synth_ragged_prompts(R, max_len=500, seed=8, groups=5, shared=120)
    seeds = np.full(R, eng.model_seed("tiny"), np.uint64)
    snap, prof, idx = o.SnapshotData(**sd), profile_of(pkg, scorers), o.Index()
    w = o.schedule_batch(snap, prof, idx, R, prompt_bytes=data, prompt_off=off, model_seed=seeds, block_chars=4,
                         max_blocks=300, want_hashes=True)
    idx.commit(w["pick"], w["hashes_out"], w["total_blocks"])
    eng.commit_picks(w["pick"], w["hashes_out"], w["total_blocks"])
    want = o.schedule_batch(snap, prof, idx, R, prompt_bytes=data, prompt_off=off, model_seed=seeds, block_chars=4,
                            max_blocks=300, want_match=True)
    got = eng.schedule(R, prompt_bytes=data, prompt_off=off, model_seed=seeds, want_match=True)
    assert_same(got, want, ("pick", "pick_score", "tie_count", "total_blocks", "match_blocks"))
    assert got["match_blocks"].max() >= 30  # 120-byte shared prefixes / 4-byte blocks
    got2 = eng.schedule(R, hashes_in=w["hashes_out"], n_hashes_in=w["total_blocks"], want_match=True)
    assert_same(got2, want, ("pick", "pick_score", "tie_count", "total_blocks", "match_blocks"))
    eng.close()


def test_device_resident_api_with_torch(pkg):
    """location=1: CUDA tensors in, CUDA tensors out, asynchronous on the caller's stream."""
    import torch
    M, R = 1024, 8192
    scorers = [("queue", 2), ("kv", 2), ("prefix", 3), ("lora", 1)]
    eng = make_engine(pkg, scorers, M, prefix_capacity=1 << 19)
    sd = synth_snapshot(M, seed=2)
    dev = torch.device("cuda:0")
    t = {k: torch.from_numpy(v).to(dev) for k, v in sd.items()}
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        eng.set_snapshot(t["kv_usage"], t["queue"], t["running"], t["lora_active"], t["lora_waiting"], t["lora_nmodels"],
                         t["lora_max"], device=True, stream=stream.cuda_stream, M=M, lora_words=1)
    prompts, off, _ = synth_prompts(R, seed=2)
    seeds = np.full(R, eng.model_seed("dev"), np.uint64)
    ad = zipf_adapters(R, seed=2)
    snap, prof, idx = o.SnapshotData(**sd), profile_of(pkg, scorers), o.Index()
    w = o.schedule_batch(snap, prof, idx, R, prompt_bytes=prompts, prompt_off=off, model_seed=seeds, adapter_id=ad,
                         want_hashes=True, n_threads=8)
    idx.commit(w["pick"], w["hashes_out"], w["total_blocks"])
    eng.commit_picks(w["pick"], w["hashes_out"], w["total_blocks"])
    want = o.schedule_batch(snap, prof, idx, R, prompt_bytes=prompts, prompt_off=off, model_seed=seeds, adapter_id=ad,
                            n_threads=8)
    out = dict(pick=torch.empty(R, dtype=torch.int32, device=dev), pick_score=torch.empty(R, dtype=torch.float64, device=dev),
               tie_count=torch.empty(R, dtype=torch.int32, device=dev), total_blocks=torch.empty(R, dtype=torch.uint16, device=dev))
    with torch.cuda.stream(stream):
        eng.schedule(R, prompt_bytes=torch.from_numpy(prompts).to(dev), prompt_off=torch.from_numpy(off).to(dev),
                     model_seed=torch.from_numpy(seeds).to(dev), adapter_id=torch.from_numpy(ad).to(dev), device=True,
                     stream=stream.cuda_stream, out=out)
    stream.synchronize()
    assert np.array_equal(out["pick"].cpu().numpy(), want["pick"])
    assert np.array_equal(out["pick_score"].cpu().numpy(), want["pick_score"])
    assert np.array_equal(out["tie_count"].cpu().numpy(), want["tie_count"])
    assert eng.stats().kernel_launches >= 3
    # the same batch cut into slices over several streams (engine debug keys 5 / 6), eagerly and inside a CUDA graph:
    # requests are independent given (snapshot, table), so every cut gives the same results
    d_in = dict(prompt_bytes=torch.from_numpy(prompts).to(dev), prompt_off=torch.from_numpy(off).to(dev),
                model_seed=torch.from_numpy(seeds).to(dev), adapter_id=torch.from_numpy(ad).to(dev))
    for split, streams in ((3, 2), (4, 3)):
        eng.set_debug(5, split)
        eng.set_debug(6, streams)
        with torch.cuda.stream(stream):
            for v in out.values():
                v.zero_()
            eng.schedule(R, device=True, stream=stream.cuda_stream, out=out, **d_in)
        stream.synchronize()
        for k in ("pick", "pick_score", "tie_count", "total_blocks"):
            assert np.array_equal(out[k].cpu().numpy(), want[k]), (split, streams, k)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        for v in out.values():
            v.zero_()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=stream):
        eng.schedule(R, device=True, stream=stream.cuda_stream, out=out, **d_in)
    g.replay()
    torch.cuda.synchronize()
    for k in ("pick", "pick_score", "tie_count", "total_blocks"):
        assert np.array_equal(out[k].cpu().numpy(), want[k]), ("graph", k)
    eng.close()


def test_full_size_headline_parity(pkg):
    """BASELINE.json headline shape: 64K requests x 1024 endpoints, all four scorers, 2 KB prompts."""
    M, R = 1024, 65536
    scorers = [("queue", 2), ("kv", 2), ("prefix", 3), ("lora", 1)]
    eng = make_engine(pkg, scorers, M, prefix_capacity=1 << 18)
    sd = synth_snapshot(M, seed=0)
    eng.set_snapshot(**sd)
    snap, prof, idx = o.SnapshotData(**sd), profile_of(pkg, scorers), o.Index()
    seed = eng.model_seed("headline")
    wp, woff, _ = synth_prompts(4 * M, seed=0)
    ws = np.full(4 * M, seed, np.uint64)
    warm = o.schedule_batch(snap, prof, idx, 4 * M, prompt_bytes=wp, prompt_off=woff, model_seed=ws, want_hashes=True,
                            n_threads=8)
    idx.commit(warm["pick"], warm["hashes_out"], warm["total_blocks"])
    eng.commit_picks(warm["pick"], warm["hashes_out"], warm["total_blocks"])
    prompts, off, _ = synth_prompts(R, seed=0)
    seeds = np.full(R, seed, np.uint64)
    ad = zipf_adapters(R, seed=0)
    got = eng.schedule(R, prompt_bytes=prompts, prompt_off=off, model_seed=seeds, adapter_id=ad, want_match=True)
    want = o.schedule_batch(snap, prof, idx, R, prompt_bytes=prompts, prompt_off=off, model_seed=seeds, adapter_id=ad,
                            want_match=True, n_threads=16)
    assert_same(got, want, ("pick", "pick_score", "tie_count", "total_blocks", "match_blocks"))
    assert (got["total_blocks"] == 32).all() and got["match_blocks"].max() == 16
    # size-independent properties: determinism, and shard-invariance (two half batches == one batch)
    again = eng.schedule(R, prompt_bytes=prompts, prompt_off=off, model_seed=seeds, adapter_id=ad)
    assert np.array_equal(again["pick"], got["pick"])
    h = R // 2
    lo = eng.schedule(h, prompt_bytes=prompts[: off[h]], prompt_off=off[: h + 1], model_seed=seeds[:h], adapter_id=ad[:h])
    hi = eng.schedule(R - h, prompt_bytes=prompts[off[h]:], prompt_off=off[h:] - off[h], model_seed=seeds[h:],
                      adapter_id=ad[h:], request_base=h)
    assert np.array_equal(np.concatenate([lo["pick"], hi["pick"]]), got["pick"])
    eng.close()


FULL_CONFIGS = [
    # BASELINE.json configs[1..4] at their full per-GPU sizes: name, M, R, scorers, prompts?, adapters?, request_base
    ("B_64K_x_256_kv_queue", 256, 65536, [("kv", 1), ("queue", 1)], False, False, 0),
    ("C_64K_x_512_prefix_2KB", 512, 65536, [("queue", 2), ("kv", 2), ("prefix", 3)], True, False, 0),
    ("D_256K_x_1024_lora_kv", 1024, 262144, [("lora", 1), ("kv", 1)], False, True, 0),
    ("E_shard_128K_x_4096_all_four", 4096, 131072, [("queue", 2), ("kv", 2), ("prefix", 3), ("lora", 1)], True, True, 3 * 131072),
]


@pytest.mark.parametrize("name,M,R,scorers,use_prompts,use_adapters,base", FULL_CONFIGS, ids=[c[0] for c in FULL_CONFIGS])
def test_full_size_baseline_configs(pkg, name, M, R, scorers, use_prompts, use_adapters, base):
    """Every request of the full-size configuration against the oracle (picks, float64 scores, tie counts, block totals);
    E is one GPU's shard of the 1M x 4096 configuration (request_base = its offset in the global batch)."""
    eng = make_engine(pkg, scorers, M, prefix_capacity=1 << 18, tie_mode=1, tie_seed=8)
    sd = synth_snapshot(M, seed=12)
    eng.set_snapshot(**sd)
    snap, prof = o.SnapshotData(**sd), profile_of(pkg, scorers, tie_mode=1, tie_seed=8)
    kw, idx = {}, None
    if use_prompts:
        idx = o.Index()
        seed = eng.model_seed("full")
        nwarm = 4 * M
        wp, woff, _ = synth_prompts(nwarm, seed=12, prefix_seed=5)
        wkw = dict(adapter_id=zipf_adapters(nwarm, seed=13)) if use_adapters else {}
        warm = o.schedule_batch(snap, prof, idx, nwarm, prompt_bytes=wp, prompt_off=woff, model_seed=np.full(nwarm, seed, np.uint64),
                                want_hashes=True, n_threads=32, **wkw)
        idx.commit(warm["pick"], warm["hashes_out"], warm["total_blocks"])
        eng.commit_picks(warm["pick"], warm["hashes_out"], warm["total_blocks"])
        prompts, off, _ = synth_prompts(R, seed=14, prefix_seed=5)
        kw.update(prompt_bytes=prompts, prompt_off=off, model_seed=np.full(R, seed, np.uint64))
    if use_adapters:
        kw["adapter_id"] = zipf_adapters(R, seed=15)
    got = eng.schedule(R, request_base=base, **kw)
    want = o.schedule_batch(snap, prof, idx, R, request_base=base, n_threads=32, **kw)
    assert_same(got, want, ("pick", "pick_score", "tie_count", "total_blocks"))
    assert (got["pick"] >= 0).all()
    if use_prompts:
        assert (got["total_blocks"] == 32).all()
    eng.close()


# ------------------------------------------------------------------------------------------ latency fold-in, token load (f2)
LAT_COEF = dict(ttft_intercept=12.5, ttft_kv=80.0, ttft_input=0.031, ttft_waiting=7.25, ttft_running=1.5,
                ttft_prefix=-40.0, tpot_intercept=9.0, tpot_kv=11.0, tpot_input=0.0007, tpot_waiting=0.9,
                tpot_running=0.35, tpot_generated=0.01)
LAT_CASES = [
    # name, M, R, scorers, latency params, options
    ("latency_only_streaming", 1024, 1024, [("latency", 1)], dict(streaming_mode=1), dict(prefix=True)),
    ("latency_defaults_nonstreaming", 300, 512, [("latency", 1)], {}, dict(prefix=True)),
    ("latency_most_masked", 512, 512, [("latency", 1)], dict(streaming_mode=1, strategy_most=1), dict(prefix=True, mask=0.3)),
    ("latency_lora_kv_buffer", 1024, 512, [("latency", 2), ("lora", 1), ("kv", 0.5)],
     dict(streaming_mode=1, slo_buffer_factor=0.9, ttft_weight=0.5, tpot_weight=1.5), dict(prefix=True, lora=True)),
    ("latency_after_queue_masked", 200, 512, [("queue", 1), ("latency", 3), ("prefix", 1)], dict(streaming_mode=1),
     dict(prefix=True, mask=0.6)),
    ("latency_composite_fallback", 700, 512, [("latency", 1)], dict(has_predictions=0, composite_kv=2.0),
     dict(prefix=True, mask=0.5)),
    ("latency_composite_zero_weights", 64, 256, [("latency", 1), ("kv", 1)],
     dict(has_predictions=0, composite_kv=0.0, composite_queue=0.0, composite_prefix=0.0), {}),
    ("latency_all_busy_buckets", 256, 512, [("latency", 1)], dict(streaming_mode=1), dict(prefix=True, busy=True)),
    ("latency_no_prefix_info_M4096", 4096, 128, [("latency", 1)], dict(streaming_mode=1, ttft_weight=0.0, tpot_weight=0.0), {}),
]


@pytest.mark.parametrize("name,M,R,scorers,lkw,opt", LAT_CASES, ids=[c[0] for c in LAT_CASES])
def test_latency_fold_in_parity(pkg, name, M, R, scorers, lkw, opt):
    """Per (request, endpoint): Bayesian-ridge TTFT/TPOT, headrooms, tier/bucket selection over the candidates,
    normalised latency score — picks, tie counts, weighted scores AND the predictions bit-exact vs the oracle."""
    import zlib
    seed = zlib.crc32(name.encode()) % 1000
    rng = np.random.Generator(np.random.PCG64(seed))
    lkw = dict(LAT_COEF, **lkw)
    eng = make_engine(pkg, scorers, M, tie_mode=1, tie_seed=5, prefix_capacity=1 << 16)
    eng.set_latency_params(pkg.latency_params(**lkw))
    sd = synth_snapshot(M, seed=seed)
    sd["min_tpot_slo"] = rng.choice([0.0, 0.0, 22.0, 26.5, 60.0], M)
    sd["dispatched"] = (rng.integers(1, 4, M) if opt.get("busy") else rng.integers(0, 3, M)).astype(np.int32)
    sd["prefill_role"] = (rng.random(M) < 0.15).astype(np.uint8)
    eng.set_snapshot(**sd)
    snap = o.SnapshotData(**sd)
    prof = o.make_profile([(pkg.SCORER[k], w) for k, w in scorers], tie_mode=1, tie_seed=5,
                          latency=o.make_latency_params(**lkw))
    kw = dict(input_tokens=rng.integers(0, 6000, R).astype(np.int32),
              ttft_slo=rng.choice([0.0, 90.0, 140.0, 200.0, 400.0, 1e6], R),
              tpot_slo=rng.choice([0.0, 18.0, 24.0, 30.0, 80.0], R))
    idx = None
    if opt.get("prefix"):
        prompts, off, _ = synth_prompts(R, prompt_len=1024, groups=12, shared=512, seed=seed)
        seeds = np.full(R, eng.model_seed("model-x"), np.uint64)
        kw.update(prompt_bytes=prompts, prompt_off=off, model_seed=seeds)
        idx = o.Index()
        warm_prof = profile_of(pkg, [("kv", 1)], tie_mode=1, tie_seed=5)
        warm = o.schedule_batch(snap, warm_prof, idx, R, prompt_bytes=prompts, prompt_off=off, model_seed=seeds,
                                want_hashes=True, n_threads=8)
        keep = np.arange(R) % 3 != 0
        idx.commit(warm["pick"][keep], warm["hashes_out"][keep], warm["total_blocks"][keep])
        eng.commit_picks(warm["pick"][keep], warm["hashes_out"][keep], warm["total_blocks"][keep])
    if opt.get("lora"):
        kw["adapter_id"] = zipf_adapters(R, seed=seed)
    if "mask" in opt:
        bits = rng.random((R, M)) < opt["mask"]
        bits[0, :] = False
        bits[1, :] = False
        bits[1, M // 2] = True  # a single candidate: zero ranges on both dimensions
        mask = np.zeros((R, (M + 31) // 32), np.uint32)
        for w in range(mask.shape[1]):
            chunk = bits[:, w * 32:(w + 1) * 32]
            mask[:, w] = (chunk * (1 << np.arange(chunk.shape[1], dtype=np.uint64))).sum(axis=1).astype(np.uint32)
        kw["cand_mask"] = mask
    got = eng.schedule(R, want_match=True, want_scores=True, want_pred=True, **kw)
    want = o.schedule_batch(snap, prof, idx, R, want_match=True, want_scores=True, want_pred=True, want_tie_set=True,
                            n_threads=8, **kw)
    assert_same(got, want, ("pick", "pick_score", "tie_count", "total_blocks", "match_blocks", "scores_out"))
    ok = want["pick"] >= 0  # (the oracle returns before predicting when a request has no candidates)
    gp, wp = got["pred_out"][ok], want["pred_out"][ok]
    assert np.array_equal(np.isnan(gp), np.isnan(wp))
    assert np.array_equal(gp[~np.isnan(gp)].view(np.uint64), wp[~np.isnan(wp)].view(np.uint64))
    fast = eng.schedule(R, **kw)   # no diagnostics: the variant without the R x M stores
    assert_same(fast, want, ("pick", "pick_score", "tie_count", "total_blocks"))
    ts = want["tie_set"]
    r = np.nonzero(got["pick"] >= 0)[0]
    assert ((ts[r, got["pick"][r] >> 5] >> (got["pick"][r] & 31).astype(np.uint32)) & 1).all()
    if lkw.get("has_predictions", 1):
        assert np.isfinite(gp).all()
        if scorers == [("latency", 1)]:  # scores are w/100 with w in [1,101] inside the chosen tier, 0 outside; clamped
            sc = got["scores_out"][~np.isnan(got["scores_out"])]
            assert sc.min() >= 0.0 and sc.max() <= 1.0 and len(np.unique(sc)) > 3
    if opt.get("prefix"):
        assert got["match_blocks"].max() > 0
    if "mask" in opt:
        assert got["pick"][0] == -1 and got["pick"][1] == M // 2
    eng.close()


def test_latency_scorer_reference_cases(pkg, golden):
    """The reference's own latency-scorer tests (plugin_test.go:52-208), driven through the engine: headrooms are
    produced by an identity-like model (TTFT = -kv_c*kv ... ) so that the scorer sees the test's headroom values."""
    for c in golden["latency_scorer"]["cases"]:
        eps = c["endpoints"]
        M = len(eps)
        eng = make_engine(pkg, [("latency", 1)], M)
        if c["info"] is None:
            eng.set_latency_params(pkg.latency_params(has_predictions=0))
            eng.set_snapshot(np.array([e[0] for e in eps]), np.array([e[1] for e in eps], np.int64),
                             np.array([e[2] for e in eps], np.int64))
        else:
            # TTFT = 1*waiting, TPOT = 1*running with SLO 1000 => headroom = 1000 - value: encode the wanted headrooms
            th = [i[0] for i in c["info"]]
            ph = [i[1] for i in c["info"]]
            eng.set_latency_params(pkg.latency_params(ttft_waiting=1.0, tpot_running=1.0, streaming_mode=1))
            eng.set_snapshot(np.zeros(M), np.array([1000 - int(t) for t in th], np.int64),
                             np.array([1000 - int(p) for p in ph], np.int64),
                             dispatched=np.array([i[2] for i in c["info"]], np.int32))
        got = eng.schedule(1, want_scores=True, ttft_slo=np.array([1000.0]), tpot_slo=np.array([1000.0]))
        sc = got["scores_out"][0]
        assert np.allclose(sc, np.clip(c["derived"], 0, 1), atol=1e-12), (c["name"], sc)
        a = c["assert"]
        for i in a.get("nonzero", []):
            assert sc[i] != 0, c["name"]
        for i in a.get("zero", []):
            assert sc[i] == 0, c["name"]
        for hi, lo in a.get("greater", []):
            assert sc[hi] > sc[lo], c["name"]
        eng.close()


def test_token_load_scorer_parity(pkg, golden):
    g = golden["token_load_scorer"]
    eng = make_engine(pkg, [("token_load", 1)], 3, token_load_threshold=g["threshold"])
    eng.set_snapshot(np.zeros(3), np.zeros(3, np.int64), inflight_tokens=np.array(g["tokens"], np.int64))
    got = eng.schedule(1, want_scores=True)
    assert np.allclose(got["scores_out"][0], g["want"], atol=g["tolerance"])
    eng.close()
    M, R = 1000, 256
    scorers = [("token_load", 1.5), ("kv", 1), ("prefix", 2)]
    rng = np.random.Generator(np.random.PCG64(3))
    sd = synth_snapshot(M, seed=3)
    sd["inflight_tokens"] = rng.choice([-3, 0, 1, 4096, 2 ** 21, 2 ** 22, 2 ** 23, 12345678], M).astype(np.int64)
    for thr, masked in ((0.0, False), (3.0e6, True)):
        eng = make_engine(pkg, scorers, M, token_load_threshold=thr)
        eng.set_snapshot(**sd)
        snap = o.SnapshotData(**sd)
        prof = o.make_profile([(pkg.SCORER[k], w) for k, w in scorers], token_load_threshold=thr)
        kw = {}
        if masked:
            kw["cand_mask"] = rng.integers(0, 2 ** 32, (R, (M + 31) // 32), dtype=np.uint64).astype(np.uint32)
        got = eng.schedule(R, want_scores=True, **kw)
        want = o.schedule_batch(snap, prof, None, R, want_scores=True, **kw)
        assert_same(got, want, ("pick", "pick_score", "tie_count", "scores_out"))
        eng.close()


# ------------------------------------------------------------------------------------------ stochastic pickers
def test_stochastic_pickers_reference_distributions(pkg, golden):
    """The reference's picker tests (weightedrandom/picker_test.go:30-140, random/picker_test.go:30-140) on the engine:
    selection frequencies within the reference's own +-5 % of score/total (resp. uniform), zero scores never picked,
    and bit-identical to the oracle, which runs the same counter-based generator."""
    g = golden["stochastic_pickers"]
    n, tol = g["iterations"], g["tolerance"]
    for mode, cases in ((pkg.PICK_WEIGHTED_RANDOM, g["weighted"]), (pkg.PICK_RANDOM, g["random"])):
        for c in cases:
            sc = np.array(c["scores"], np.float64)
            M = len(sc)
            eng = make_engine(pkg, [("col0", 100.0)], M, pick_mode=mode, tie_seed=1234)
            eng.set_snapshot(np.zeros(M), np.zeros(M, np.int64), endpoint_cols=[sc / 100.0])
            got = eng.schedule(n)
            snap = o.SnapshotData(np.zeros(M), np.zeros(M, np.int64), endpoint_cols=[sc / 100.0])
            want = o.schedule_batch(snap, o.make_profile([(o.SCORER_ENDPOINT_COL0, 100.0)], tie_seed=1234, pick_mode=mode), None, n)
            assert_same(got, want)
            freq = np.bincount(got["pick"], minlength=M) / n
            expect = sc / sc.sum() if mode == pkg.PICK_WEIGHTED_RANDOM else np.full(M, 1.0 / M)
            assert np.abs(freq - expect).max() <= tol, (c["name"], freq)
            if mode == pkg.PICK_WEIGHTED_RANDOM:
                assert (freq[sc == 0] == 0).all()
            eng.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_stochastic_pickers_parity_full_profile(pkg, mode):
    """All four scorers + candidate masks under the stochastic pickers: pick, score of the pick and the size of the
    draw set equal the oracle's; request_base shards agree (two half batches == one batch)."""
    M, R = 600, 2048
    scorers = [("queue", 2), ("kv", 2), ("prefix", 3), ("lora", 1)]
    eng = make_engine(pkg, scorers, M, pick_mode=mode, tie_seed=77, prefix_capacity=1 << 16)
    sd = synth_snapshot(M, seed=11)
    eng.set_snapshot(**sd)
    snap = o.SnapshotData(**sd)
    prof = o.make_profile([(pkg.SCORER[k], w) for k, w in scorers], tie_seed=77, pick_mode=mode)
    prompts, off, _ = synth_prompts(R, prompt_len=1024, groups=10, shared=512, seed=3)
    seeds = np.full(R, eng.model_seed("m"), np.uint64)
    idx = o.Index()
    warm = o.schedule_batch(snap, profile_of(pkg, [("kv", 1)]), idx, R, prompt_bytes=prompts, prompt_off=off,
                            model_seed=seeds, want_hashes=True, n_threads=8)
    idx.commit(warm["pick"], warm["hashes_out"], warm["total_blocks"])
    eng.commit_picks(warm["pick"], warm["hashes_out"], warm["total_blocks"])
    rng = np.random.Generator(np.random.PCG64(5))
    mask = rng.integers(0, 2 ** 32, (R, (M + 31) // 32), dtype=np.uint64).astype(np.uint32)
    mask[0, :] = 0
    kw = dict(prompt_bytes=prompts, prompt_off=off, model_seed=seeds, adapter_id=zipf_adapters(R, seed=3))
    for extra in ({}, {"cand_mask": mask}):
        got = eng.schedule(R, want_scores=True, **kw, **extra)
        want = o.schedule_batch(snap, prof, idx, R, want_scores=True, want_tie_set=True, n_threads=8, **kw, **extra)
        assert_same(got, want, ("pick", "pick_score", "tie_count", "scores_out"))
        fast = eng.schedule(R, **kw, **extra)
        assert_same(fast, want)
    assert got["pick"][0] == -1
    assert len(np.unique(got["pick"])) > 50   # spread over many endpoints, unlike the arg-max
    h = R // 2
    a = eng.schedule(h, prompt_bytes=prompts[: off[h]], prompt_off=off[: h + 1], model_seed=seeds[:h], adapter_id=kw["adapter_id"][:h])
    b = eng.schedule(R - h, prompt_bytes=prompts[off[h]:], prompt_off=off[h:] - off[h], model_seed=seeds[h:],
                     adapter_id=kw["adapter_id"][h:], request_base=h)
    full = eng.schedule(R, **kw)
    assert np.array_equal(np.concatenate([a["pick"], b["pick"]]), full["pick"])
    eng.close()


def test_masked_minmax_bucket_and_scan_paths(pkg):
    """Masked queue / running scorers: the value-bucket fast path (range <= 255), the scanning fallback (wide range) and a
    mix of both in one profile; sparse masks, single-candidate and empty rows."""
    M, R = 777, 1024
    scorers = [("queue", 2), ("running", 1.5), ("kv", 1)]
    rng = np.random.Generator(np.random.PCG64(21))
    for qmax, rmax in ((40, 12), (10 ** 6, 12), (40, 10 ** 9), (3, 1), (0, 0)):
        sd = synth_snapshot(M, seed=9)
        sd["queue"] = rng.integers(0, qmax + 1, M).astype(np.int64) - 5
        sd["running"] = rng.integers(0, rmax + 1, M).astype(np.int64)
        eng = make_engine(pkg, scorers, M, tie_mode=1, tie_seed=3)
        eng.set_snapshot(**sd)
        snap = o.SnapshotData(**sd)
        prof = profile_of(pkg, scorers, tie_mode=1, tie_seed=3)
        bits = rng.random((R, M)) < rng.choice([0.01, 0.1, 0.5, 1.0], (R, 1))
        bits[0, :] = False
        bits[1, :] = False
        bits[1, 700] = True
        mask = np.zeros((R, (M + 31) // 32), np.uint32)
        for w in range(mask.shape[1]):
            chunk = bits[:, w * 32:(w + 1) * 32]
            mask[:, w] = (chunk * (1 << np.arange(chunk.shape[1], dtype=np.uint64))).sum(axis=1).astype(np.uint32)
        mask[:, -1] |= np.uint32((0xFFFFFFFF << (M % 32)) & 0xFFFFFFFF)  # garbage bits beyond M must be ignored
        got = eng.schedule(R, cand_mask=mask, want_scores=True)
        want = o.schedule_batch(snap, prof, None, R, cand_mask=mask, want_scores=True, n_threads=8)
        # the oracle reads only bits < M; rows 0 and 1 are the empty / single-candidate rows
        assert_same(got, want, ("pick", "pick_score", "tie_count", "scores_out"))
        fast = eng.schedule(R, cand_mask=mask)
        assert_same(fast, want)
        assert got["pick"][0] == -1 and got["pick"][1] == 700
        eng.close()


def _texty_prompts(R, seed, max_len=1400):
    """Ragged prompts made of words, ASCII / Unicode spaces and some invalid UTF-8; returns bytes, offsets."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pieces = [b"the", b"quick", b"brown", b"fox", b"\xe7\x8c\xab", b"x", b"", b"\xf0\x9f\x98\x80", b"\xff", b"\xe2\x80", b"\xc2"]
    spaces = [b" ", b" ", b" ", b"\t", b"\n", b"  ", b"\xc2\xa0", b"\xe2\x80\x83", b"\xe3\x80\x80", b"\xc2\x85", b"\xe2\x80\xa8",
              b"\xe1\x9a\x80", b"\xe2\x81\x9f", b"\r\n"]
    out, off = bytearray(), [0]
    for r in range(R):
        target = 0 if r == 0 else int(rng.integers(0, max_len))
        s = bytearray()
        while len(s) < target:
            s += pieces[int(rng.integers(0, len(pieces)))]
            s += spaces[int(rng.integers(0, len(spaces)))]
        out += s[:target]
        off.append(len(out))
    return np.frombuffer(bytes(out) + b"\0" * 64, np.uint8), np.array(off, np.int64)


def test_count_fields_parity(pkg):
    """len(strings.Fields(prompt)) on the device vs Go's semantics restated in the oracle: ragged prompts, Unicode spaces,
    invalid UTF-8, prompts crossing the 512-byte step, back-to-back (unaligned) and 16-byte aligned starts."""
    eng = make_engine(pkg, [("kv", 1)], 8)
    for seed in (1, 2):
        data, off = _texty_prompts(600, seed)
        R = len(off) - 1
        want = np.array([o.count_fields(bytes(data[off[r]:off[r + 1]])) for r in range(R)], np.int32)
        assert np.array_equal(eng.count_fields(data, off), want)
        # 16-byte aligned starts with explicit lengths
        buf, aoff, lens = bytearray(), [], []
        for r in range(R):
            buf.extend(b"\x20" * ((-len(buf)) % 16))
            aoff.append(len(buf))
            lens.append(int(off[r + 1] - off[r]))
            buf.extend(bytes(data[off[r]:off[r + 1]]))
        aoff.append(len(buf))
        got = eng.count_fields(np.frombuffer(bytes(buf) + b"\0" * 64, np.uint8), np.array(aoff, np.int64), np.array(lens, np.int32))
        assert np.array_equal(got, want)
    assert want.max() > 100 and (want == 0).any()
    eng.close()


def test_latency_counts_prompt_fields_on_device(pkg):
    """input_tokens omitted: the engine counts the fields of the prompt bytes itself (predictedlatency/plugin.go:286)."""
    M, R = 200, 300
    lkw = dict(LAT_COEF, streaming_mode=1)
    eng = make_engine(pkg, [("latency", 1)], M, tie_mode=1, tie_seed=9)
    eng.set_latency_params(pkg.latency_params(**lkw))
    sd = synth_snapshot(M, seed=4)
    eng.set_snapshot(**sd)
    data, off = _texty_prompts(R, 5, max_len=3000)
    seeds = np.full(R, eng.model_seed("m"), np.uint64)
    slo = dict(ttft_slo=np.full(R, 180.0), tpot_slo=np.full(R, 28.0))
    got = eng.schedule(R, prompt_bytes=data, prompt_off=off, model_seed=seeds, want_scores=True, **slo)
    toks = np.array([o.count_fields(bytes(data[off[r]:off[r + 1]])) for r in range(R)], np.int32)
    prof = o.make_profile([(o.SCORER_LATENCY, 1.0)], tie_mode=1, tie_seed=9, latency=o.make_latency_params(**lkw))
    want = o.schedule_batch(o.SnapshotData(**sd), prof, o.Index(), R, prompt_bytes=data, prompt_off=off, model_seed=seeds,
                            input_tokens=toks, want_scores=True, **slo)
    assert_same(got, want, ("pick", "pick_score", "tie_count", "scores_out"))
    assert toks.max() > 300
    eng.close()


# ------------------------------------------------------------------------------------------ device-side filters
def test_filters_reference_cases_on_device(pkg, golden):
    """The reference's filter tests (sloheadroomtier/plugin_test.go:43-115, prefixcacheaffinity/plugin_test.go:48-105) through
    the engine: the surviving candidate set is exactly the one the Go tests expect."""
    from tests.test_oracle_golden import filter_case_inputs
    for kind, key in (("tier", "slo_headroom_tier"), ("affinity", "prefix_cache_affinity")):
        for c in golden["filters"][key]["cases"]:
            inp = filter_case_inputs(kind, c)
            M = inp["M"]
            filters = [(k, par) for k, par in inp["filters"]]
            eng = pkg.Engine(pkg.default_config([("latency", 1.0)], filters=filters, max_endpoints=M, max_blocks=128, tie_seed=3,
                                                prefix_capacity=1 << 12))
            eng.set_latency_params(pkg.latency_params(**inp["lat"]))
            eng.set_snapshot(np.zeros(M), np.array(inp["queue"], np.int64), np.array(inp["running"], np.int64))
            for m, n in inp["adds"]:
                eng.prefix_add(inp["hashes"][:n], m)
            res = eng.schedule(1, hashes_in=inp["hashes"][None, :].copy(), n_hashes_in=np.array([100], np.uint16),
                               ttft_slo=np.array([1000.0]), tpot_slo=np.array([1000.0]), want_filter_mask=True)
            kept = [m for m in range(M) if (res["filter_mask_out"][0][m >> 5] >> (m & 31)) & 1]
            assert kept == c["want"], (c["name"], kept)
            assert res["pick"][0] in kept
            eng.close()


@pytest.mark.parametrize("masked", [False, True])
def test_latency_chart_profile_parity(pkg, masked):
    """The latency profile of the reference chart (config/charts/epplib/templates/_config.yaml:66-75) entirely on the device:
    strict affinity filter -> slo-headroom-tier filter -> loose affinity filter -> latency scorer -> weighted-random picker.
    Exploration probabilities are raised so that every branch is taken; filter results, picks, scores: bit-equal to the oracle."""
    M, R = 300, 1024
    lkw = dict(LAT_COEF, streaming_mode=1)
    filt = [(pkg.FILTER_PREFIX_AFFINITY, (0.95, 0.2, 40.0)), (pkg.FILTER_SLO_HEADROOM_TIER, (0.3,)),
            (pkg.FILTER_PREFIX_AFFINITY, (0.50, 0.1, 5000.0))]
    scorers = [("latency", 1.0)]
    eng = pkg.Engine(pkg.default_config(scorers, filters=filt, max_endpoints=M, max_blocks=16, pick_mode=pkg.PICK_WEIGHTED_RANDOM, tie_seed=21,
                                        prefix_capacity=1 << 15))
    eng.set_latency_params(pkg.latency_params(**lkw))
    rng = np.random.Generator(np.random.PCG64(31))
    sd = synth_snapshot(M, seed=6)
    sd["min_tpot_slo"] = rng.choice([0.0, 22.0, 60.0], M)
    sd["dispatched"] = rng.integers(0, 3, M).astype(np.int32)
    sd["prefill_role"] = (rng.random(M) < 0.1).astype(np.uint8)
    eng.set_snapshot(**sd)
    snap = o.SnapshotData(**sd)
    prof = o.make_profile([(o.SCORER_LATENCY, 1.0)], latency=o.make_latency_params(**lkw), pick_mode=o.PICK_WEIGHTED_RANDOM, tie_seed=21,
                          filters=[(k, par) for k, par in filt])
    prompts, off, _ = synth_prompts(R, prompt_len=1024, groups=6, shared=1024 - 64, seed=8)  # 15 of 16 blocks shared: scores up to 1.0
    seeds = np.full(R, eng.model_seed("m"), np.uint64)
    idx = o.Index()
    # warm-up: every request lands on a uniformly random endpoint, so each shared prefix is cached on many endpoints;
    # half of the warm-up prompts are truncated to 8 blocks so that partial matches (score 0.5) exist as well
    warm_prof = o.make_profile([(o.SCORER_KV_CACHE, 1.0)], pick_mode=o.PICK_RANDOM, tie_seed=2)
    warm = o.schedule_batch(snap, warm_prof, idx, R, prompt_bytes=prompts, prompt_off=off, model_seed=seeds, max_blocks=16,
                            want_hashes=True, n_threads=8)
    nblk = np.where(np.arange(R) % 2 == 0, warm["total_blocks"], np.minimum(warm["total_blocks"], 8)).astype(np.uint16)
    idx.commit(warm["pick"], warm["hashes_out"], nblk)
    eng.commit_picks(warm["pick"], warm["hashes_out"], nblk)
    kw = dict(prompt_bytes=prompts, prompt_off=off, model_seed=seeds, input_tokens=rng.integers(0, 3000, R).astype(np.int32),
              ttft_slo=rng.choice([0.0, 120.0, 180.0, 400.0], R), tpot_slo=rng.choice([0.0, 22.0, 30.0], R))
    if masked:
        mask = rng.integers(0, 2 ** 32, (R, (M + 31) // 32), dtype=np.uint64).astype(np.uint32)
        mask[0, :] = 0
        mask[1, :] = 0
        mask[1, 3] = 1 << 7
        kw["cand_mask"] = mask
    got = eng.schedule(R, want_scores=True, want_filter_mask=True, want_match=True, **kw)
    want = o.schedule_batch(snap, prof, idx, R, max_blocks=16, want_scores=True, want_filter_mask=True, want_match=True, n_threads=8, **kw)
    lastw = np.uint32((1 << (M % 32)) - 1) if M % 32 else np.uint32(0xFFFFFFFF)
    gm, wm = got["filter_mask_out"].copy(), want["filter_mask_out"].copy()
    gm[:, -1] &= lastw
    wm[:, -1] &= lastw
    assert np.array_equal(got["match_blocks"], want["match_blocks"])
    assert np.array_equal(gm, wm)
    assert_same(got, want, ("pick", "pick_score", "tie_count"))
    # the weighted scores of the surviving candidates (the oracle reports NaN outside the filtered set, the engine outside the input mask)
    keep = np.zeros((R, M), bool)
    for w in range(gm.shape[1]):
        bits = (gm[:, w:w + 1] >> np.arange(32, dtype=np.uint32)) & 1
        keep[:, w * 32:(w + 1) * 32] = bits[:, : min(32, M - w * 32)].astype(bool)
    assert np.array_equal(got["scores_out"][keep].view(np.uint64), want["weighted_out"][keep].view(np.uint64))
    sizes = keep.sum(axis=1)
    full = (~np.isnan(want["weighted_out"]) | keep).sum(axis=1)
    assert (sizes[2:] >= 1).all() and (sizes < (M if not masked else M)).any()   # the filters did narrow some requests
    fast = eng.schedule(R, **kw)
    assert_same(fast, want, ("pick", "pick_score", "tie_count"))
    eng.close()


# ------------------------------------------------------------------------------------------ device-resident prefix index
def _lru_equal(eng, idx, eps):
    for m in eps:
        assert eng.prefix_lru_keys(m) == idx.lru_keys(m), m


def test_device_commit_closed_loop_matches_oracle(pkg):
    """schedule -> commit -> schedule entirely on the device (eppscore_commit_picks_device on the pick / hashes_out buffers of
    a device-location batch, nothing crosses PCIe) against the oracle's scheduler + indexer: picks, scores, tie counts,
    match counts every round; LRU contents (oldest -> newest) and len(hashToPods) at the end.  LRU capacity 600 with
    ~430 new blocks per endpoint per round: evictions reach back over earlier rounds, logs compact, the table is rebuilt."""
    import torch
    M, R, rounds = 96, 6000, 5
    scorers = [("queue", 2), ("kv", 2), ("prefix", 3), ("lora", 1)]
    eng = make_engine(pkg, scorers, M, lru_capacity_default=600, max_blocks=16, prefix_capacity=1 << 10)
    sd = synth_snapshot(M, seed=31)
    eng.set_snapshot(**sd)
    snap, prof, idx = o.SnapshotData(**sd), profile_of(pkg, scorers), o.Index(600)
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(device=dev)
    out = dict(pick=torch.empty(R, dtype=torch.int32, device=dev), pick_score=torch.empty(R, dtype=torch.float64, device=dev),
               tie_count=torch.empty(R, dtype=torch.int32, device=dev), total_blocks=torch.empty(R, dtype=torch.uint16, device=dev),
               hashes_out=torch.zeros((R, 16), dtype=torch.uint64, device=dev))
    for rnd in range(rounds):
        prompts, off, _ = synth_prompts(R, prompt_len=640, groups=40, shared=320, seed=40 + rnd, prefix_seed=3)
        seeds = np.full(R, eng.model_seed("loop"), np.uint64)
        ad = zipf_adapters(R, seed=rnd)
        with torch.cuda.stream(st):
            dp, do = torch.from_numpy(prompts).to(dev), torch.from_numpy(off).to(dev)
            ds, da = torch.from_numpy(seeds.view(np.int64)).to(dev), torch.from_numpy(ad).to(dev)
            eng.schedule(R, prompt_bytes=dp, prompt_off=do, model_seed=ds, adapter_id=da, request_base=rnd * R, device=True,
                         stream=st.cuda_stream, out=out)
            eng.commit_picks_device(out["pick"], out["hashes_out"], out["total_blocks"], touch_bound=R * 10, stream=st.cuda_stream)
        want = o.schedule_batch(snap, prof, idx, R, prompt_bytes=prompts, prompt_off=off, model_seed=seeds, adapter_id=ad,
                                max_blocks=16, want_hashes=True, request_base=rnd * R, n_threads=8)
        st.synchronize()
        got = {k: v.cpu().numpy() for k, v in out.items()}
        assert_same(got, want, ("pick", "pick_score", "tie_count", "total_blocks"))
        idx.commit(want["pick"], want["hashes_out"], want["total_blocks"])
        if rnd > 0:
            assert int((want["pick_score"] > 4.0).sum()) > 0  # prefix matches really contribute
    _lru_equal(eng, idx, range(M))
    s = eng.stats()
    assert s.prefix_live_hashes == idx.num_hashes() and s.index_error == 0
    assert s.lru_entries == sum(max(idx.lru_len(m), 0) for m in range(M))
    assert s.prefix_rebuilds >= 1 and s.prefix_hashes <= s.prefix_capacity
    # the table answers Get like the oracle for a sample of committed hashes
    for h in want["hashes_out"][::97, :4].reshape(-1):
        assert eng.prefix_get(int(h)) == idx.get(int(h))
    eng.close()


def test_replicas_replay_the_commit_stream(pkg):
    """Replication across GPUs is replay: two engines fed the same ordered commits hold the same index — checked at
    the headline shape (64K requests x 1024 endpoints, 2 KB prompts): engine B never sees engine A's table, only the commit stream
    (with the second batch split differently), and must return the oracle's picks, scores, tie counts and match counts."""
    M, R = 1024, 65536
    scorers = [("queue", 2), ("kv", 2), ("prefix", 3), ("lora", 1)]
    sd = synth_snapshot(M, seed=2)
    snap, prof, idx = o.SnapshotData(**sd), profile_of(pkg, scorers), o.Index()
    engs = [make_engine(pkg, scorers, M, prefix_capacity=1 << 12) for _ in range(2)]
    for e in engs:
        e.set_snapshot(**sd)
    seed = engs[0].model_seed("replica")
    wp, woff, _ = synth_prompts(4 * M, groups=150, seed=9, prefix_seed=1)
    warm = o.schedule_batch(snap, prof, idx, 4 * M, prompt_bytes=wp, prompt_off=woff, model_seed=np.full(4 * M, seed, np.uint64),
                            want_hashes=True, n_threads=8)
    idx.commit(warm["pick"], warm["hashes_out"], warm["total_blocks"])
    engs[0].commit_picks(warm["pick"], warm["hashes_out"], warm["total_blocks"])
    half = 2 * M
    engs[1].commit_picks(warm["pick"][:half], warm["hashes_out"][:half], warm["total_blocks"][:half])   # same stream,
    engs[1].commit_picks(warm["pick"][half:], warm["hashes_out"][half:], warm["total_blocks"][half:])   # cut elsewhere
    prompts, off, _ = synth_prompts(R, groups=150, seed=10, prefix_seed=1)
    kw = dict(prompt_bytes=prompts, prompt_off=off, model_seed=np.full(R, seed, np.uint64), adapter_id=zipf_adapters(R, seed=4))
    want = o.schedule_batch(snap, prof, idx, R, n_threads=8, want_hashes=True, **kw)
    for e in engs:
        got = e.schedule(R, want_hashes=True, **kw)
        assert_same(got, want, ("pick", "pick_score", "tie_count", "total_blocks"))
        e.commit_picks(got["pick"], got["hashes_out"], got["total_blocks"])
    idx.commit(want["pick"], want["hashes_out"], want["total_blocks"])
    n = 4096
    kw2 = {k: (v[: off[n]] if k == "prompt_bytes" else v[: n + 1] if k == "prompt_off" else v[:n]) for k, v in kw.items()}
    want2 = o.schedule_batch(snap, prof, idx, n, want_match=True, n_threads=8, **kw2)
    for e in engs:
        got2 = e.schedule(n, want_match=True, **kw2)
        assert_same(got2, want2, ("pick", "pick_score", "tie_count", "match_blocks"))
        assert e.stats().prefix_live_hashes == idx.num_hashes()
        e.close()
    assert int(want2["match_blocks"].max()) >= 16


def test_sets_beyond_eight_endpoints_and_universal_prefix(pkg):
    """Blocks cached on more than 10 endpoints live in bitset rows; a system prompt cached on EVERY endpoint makes every endpoint
    an exception of the sparse pick.  One batch builds those sets concurrently (every endpoint's CTA adds itself to the same
    16 slots); the sparse and the full-matrix kernels must both agree with the oracle afterwards."""
    M, R = 200, 3000
    scorers = [("queue", 2), ("kv", 2), ("prefix", 3)]
    eng = make_engine(pkg, scorers, M, tie_mode=1, tie_seed=5)
    sd = synth_snapshot(M, seed=8, tie_heavy=True)
    eng.set_snapshot(**sd)
    snap, prof, idx = o.SnapshotData(**sd), profile_of(pkg, scorers, tie_mode=1, tie_seed=5), o.Index()
    prompts, off, _ = synth_prompts(R, prompt_len=1536, groups=3, shared=1024, seed=12)  # 3 system prompts, 16 + 8 blocks
    seeds = np.full(R, eng.model_seed("sys"), np.uint64)
    hashes, nh = eng.hash_prompts(prompts, off, seeds)
    pick = (np.arange(R) % M).astype(np.int32)            # round-robin: every endpoint caches all three system prompts
    eng.commit_picks(pick, hashes, nh)
    idx.commit(pick, hashes, nh)
    st0 = eng.stats()
    distinct = len({int(h) for r in range(R) for h in hashes[r, : nh[r]]})
    bad = [(g, b, sorted(idx.get(int(hashes[g, b])) - eng.prefix_get(int(hashes[g, b]))))
           for g in range(12) for b in range(16) if eng.prefix_get(int(hashes[g, b])) != idx.get(int(hashes[g, b]))]
    assert not bad and st0.prefix_hashes == distinct == st0.prefix_live_hashes and st0.index_error == 0, \
        (bad[:6], st0.prefix_hashes, st0.prefix_live_hashes, distinct, st0.index_error)
    assert len(idx.get(int(hashes[0, 0]))) >= M - 8       # (nearly) every endpoint holds the first system prompt
    assert eng.stats().prefix_overflow_rows == 3 * 16 and eng.stats().index_error == 0
    kw = dict(prompt_bytes=prompts, prompt_off=off, model_seed=seeds)
    want = o.schedule_batch(snap, prof, idx, R, want_match=True, n_threads=8, **kw)
    got = eng.schedule(R, **kw)                                        # sparse kernel
    assert_same(got, want, ("pick", "pick_score", "tie_count"))
    full = eng.schedule(R, want_match=True, **kw)                      # full-matrix kernel
    assert_same(full, want, ("pick", "pick_score", "tie_count", "match_blocks"))
    assert int((want["match_blocks"] >= 16).sum(axis=1).min()) >= M - 8  # (nearly) every endpoint is an exception of every request
    for m in range(0, M, 2):
        eng.prefix_remove_endpoint(m)
        idx.remove_pod(m)
    want = o.schedule_batch(snap, prof, idx, R, want_match=True, n_threads=8, **kw)
    got = eng.schedule(R, want_match=True, **kw)
    assert_same(got, want, ("pick", "pick_score", "tie_count", "match_blocks"))
    assert_same(eng.schedule(R, **kw), want, ("pick", "pick_score", "tie_count"))
    eng.close()


def test_index_churn_stays_bounded_on_device(pkg):
    """emptied keys are reclaimed.  4 endpoints x LRU 100, 64 fresh hashes per Add, 1200 Adds: 76 800 hashes
    pass through, 400 stay — the table must be rebuilt in place, never grown, and stay exact."""
    eng = make_engine(pkg, [("prefix", 1.0)], 4, lru_capacity_default=100, prefix_capacity=512)
    idx = o.Index(100)
    rng = np.random.Generator(np.random.PCG64(9))
    R = 48
    for it in range(25):
        hashes = rng.integers(1, 2 ** 63, size=(R, 64), dtype=np.uint64)
        pick = (np.arange(R) % 4).astype(np.int32)
        nh = np.full(R, 64, np.uint16)
        eng.commit_picks(pick, hashes, nh)
        idx.commit(pick, hashes, nh)
    s = eng.stats()
    assert s.prefix_live_hashes == idx.num_hashes() == 400 and s.index_error == 0
    assert s.prefix_rebuilds >= 3 and s.prefix_capacity <= 8192
    _lru_equal(eng, idx, range(4))
    for h in hashes[-1, ::9]:
        assert eng.prefix_get(int(h)) == idx.get(int(h))
    eng.close()
