"""Pins the CPU oracle (oracle/oracle.c) against every golden vector the reference's own tests hold for
the hot path (tests/golden/reference_vectors.json, transcribed with file:line citations) and against
two independent XXH64 implementations (tests/golden/xxh64_kat.json).  CPU-only.
"""
import numpy as np
import pytest

from oracle import oracle_py as o
from tests.helpers import kinds, mask_from_list, pack_lora


# ---------------------------------------------------------------- XXH64 / hashPrompt
def test_xxh64_spec_and_raw_vectors(xxh_kat):
    for v in xxh_kat["spec_vectors"]:
        assert f"{o.xxh64(v['ascii'].encode()):016x}" == v["xxh64"]
    for v in xxh_kat["raw"]:
        assert f"{o.xxh64(bytes.fromhex(v['hex'])):016x}" == v["xxh64"]


def test_hash_chain_vectors(xxh_kat):
    for c in xxh_kat["chains"]:
        seed = o.model_seed(c["model"], c["salt"])
        assert f"{seed:016x}" == c["seed"]
        got = o.hash_prompt(bytes.fromhex(c["prompt_hex"]), seed, c["block_chars"], c["max_blocks"])
        assert [f"{int(x):016x}" for x in got] == c["hashes"], c


def test_hash_counts_reference(golden):
    for c in golden["hash_counts"]["cases"]:
        prompt = c["prompt"] if "prompt" in c else c["prompt_repeat"][0] * c["prompt_repeat"][1]
        got = o.hash_prompt(prompt.encode(), o.model_seed(c["model"]), c["block_chars"], c["max_blocks"])
        assert len(got) == c["want_n"], c["source"]


def test_hash_prompt_edge_cases():
    seed = o.model_seed("m")
    assert len(o.hash_prompt(b"", seed, 64, 256)) == 0
    assert len(o.hash_prompt(b"x" * 63, seed, 64, 256)) == 0          # hashing.go:57-60
    assert len(o.hash_prompt(b"x" * 64, seed, 64, 256)) == 1
    assert len(o.hash_prompt(b"x" * 65, seed, 64, 256)) == 2          # trailing partial block (:89-95)
    assert len(o.hash_prompt(b"x" * 6400, seed, 64, 16)) == 16        # truncation (:62-65)
    assert len(o.hash_prompt(b"x" * 64, seed, 0, 256)) == 0           # block size must be positive (:51-56)
    assert len(o.hash_prompt(b"x" * 64, seed, 64, 0)) == 0            # maxBlocks 0 truncates to nothing
    # different model ⇒ different chain for the same body (:69-71)
    a = o.hash_prompt(b"x" * 128, o.model_seed("m1"), 64, 256)
    b = o.hash_prompt(b"x" * 128, o.model_seed("m2"), 64, 256)
    assert a[0] != b[0] and a[1] != b[1]
    # cache salt participates in the seed (:72-75)
    assert o.model_seed("m1", "s") != o.model_seed("m1") and o.model_seed("m1", "s") == o.xxh64(b"m1s")


# ---------------------------------------------------------------- scorers
def _snap_from_endpoints(eps, target_models=()):
    ids, act, wai, nm, mx = pack_lora(eps, target_models)
    snap = o.SnapshotData(kv_usage=[e.get("kv", 0.0) for e in eps], queue=[e.get("queue", 0) for e in eps],
                          lora_active=act, lora_waiting=wai, lora_nmodels=nm, lora_max=mx)
    return snap, ids


def test_kv_scorer(golden):
    g = golden["kv_scorer"]
    for c in g["cases"]:
        snap = o.SnapshotData(kv_usage=c["kv"], queue=np.zeros(len(c["kv"])))
        got = o.score_single("kv", snap)
        assert np.allclose(got, c["want"], atol=g["tolerance"], rtol=0)


def test_queue_scorer(golden):
    g = golden["queue_scorer"]
    for c in g["cases"]:
        snap = o.SnapshotData(kv_usage=np.zeros(len(c["queue"])), queue=c["queue"])
        got = o.score_single("queue", snap)
        assert np.allclose(got, c["want"], atol=g["tolerance"], rtol=0)
    # the min/max are over the GIVEN endpoints (queue.go:79-91): a mask changes the normalisation
    snap = o.SnapshotData(kv_usage=np.zeros(3), queue=[10, 5, 0])
    got = o.score_single("queue", snap, mask=mask_from_list(3, [0, 1]))
    assert got[0] == 0.0 and got[1] == 1.0 and np.isnan(got[2])


def test_running_requests_scorer(golden):
    g = golden["running_scorer"]
    for c in g["cases"]:
        n = len(c["running"])
        snap = o.SnapshotData(kv_usage=np.zeros(n), queue=np.zeros(n, np.int64), running=c["running"])
        assert np.allclose(o.score_single("running", snap), c["want"], atol=g["tolerance"], rtol=0)


def test_lora_scorer(golden):
    g = golden["lora_scorer"]
    for c in g["cases"]:
        if not c["endpoints"]:
            continue
        snap, ids = _snap_from_endpoints(c["endpoints"], [c["target"]])
        got = o.score_single("lora", snap, adapter_id=ids[c["target"]])
        assert np.allclose(got, c["want"], atol=g["tolerance"], rtol=0), c["name"]
    # an adapter outside the dictionary is never active/waiting; the capacity rule still applies
    snap, _ = _snap_from_endpoints([{"active": ["a"], "waiting": [], "max_active": 2},
                                    {"active": ["a", "b"], "waiting": [], "max_active": 2}])
    assert list(o.score_single("lora", snap, adapter_id=-1)) == [0.8, 0.0]


def test_prefix_scorer(golden):
    for c in golden["prefix_scorer"]["cases"]:
        got = o.score_prefix(c["match"], c["total"], len(c["match"]))
        assert list(got) == c["want_exact"]                       # assert.Equal in the reference: exact
    assert list(o.score_prefix([3, 1], 0, 2)) == [0.0, 0.0]       # total==0 ⇒ 0 (plugin.go:108)
    assert list(o.score_prefix([3, 1], 4, 2, have_info=False)) == [0.0, 0.0]  # attribute absent ⇒ 0 (:100-106)


def test_enforce_score_range(golden):
    for x, want in golden["enforce_score_range"]["cases"]:
        assert o.lib().orc_enforce_score_range(x) == want
    r = golden["enforce_score_range"]["run_out_of_range"]
    snap = o.SnapshotData(kv_usage=[0.0], queue=[0], endpoint_cols=[[r["scores"][0]], [r["scores"][1]]])
    prof = o.make_profile([(8, r["weights"][0]), (9, r["weights"][1])])
    assert o.schedule_one(snap, prof)["score"] == r["want_score_exact"]


# ---------------------------------------------------------------- SchedulerProfile.Run / Scheduler.Schedule
def test_schedule_finds_optimal_endpoint(golden):
    for c in golden["schedule"]:
        eps = c["endpoints"]
        if c.get("want_error"):
            snap = o.SnapshotData(kv_usage=np.zeros(0), queue=np.zeros(0))
            res = o.schedule_one(snap, o.make_profile(kinds(c["scorers"])))
            assert res["rc"] == -1 and res["pick"] == -1
            continue
        snap, ids = _snap_from_endpoints(eps, [c["target_model"]])
        res = o.schedule_one(snap, o.make_profile(kinds(c["scorers"])), adapter_id=ids[c["target_model"]])
        assert eps[res["pick"]]["name"] == c["want_pick"]
        assert res["score"] == c["want_score_exact"]              # `==` on float64 in the reference (types.go:148-150)
        assert res["tie_count"] == 1


def test_weighted_constant_scorers(golden):
    for c in golden["weighted_constant_scorers"]:
        n = c["n_endpoints"]
        cols = [[s] * n for s in c["scores"]]
        snap = o.SnapshotData(kv_usage=np.zeros(n), queue=np.zeros(n), endpoint_cols=cols)
        prof = o.make_profile([(8 + i, w) for i, w in enumerate(c["weights"])])
        res = o.schedule_one(snap, prof, mask=mask_from_list(n, c["filter_keep"]))
        if c.get("want_error"):
            assert res["rc"] == -1 and res["pick"] == -1 and res["tie_count"] == 0
        else:
            assert res["score"] == c["want_score_exact"]
            assert res["tie_count"] == c["want_tie_count"] and res["pick"] in c["filter_keep"]
            assert res["tie_set"] == c["filter_keep"]


def test_integration_routing(golden):
    g = golden["integration_routing"]
    for c in g["cases"]:
        scorers = c.get("scorers", g["scorers"])
        eps = [{"queue": q, "kv": kv, "active": models, "waiting": [], "max_active": 0} for _, q, kv, models in c["pods"]]
        snap, ids = _snap_from_endpoints(eps, [c["target_model"]])
        idx = o.Index()
        hashes = o.hash_prompt(c["prompt"].encode(), o.model_seed(c["target_model"]), 64, 256)
        match = idx.match(hashes, len(eps))
        mask = mask_from_list(len(eps), c["subset"]) if "subset" in c else None
        res = o.schedule_one(snap, o.make_profile(kinds(scorers)), adapter_id=ids[c["target_model"]], mask=mask,
                             match=match, total=len(hashes))
        if c.get("want_error"):
            assert res["rc"] == -1, c["name"]
        else:
            assert res["pick"] == c["want_pick"], c["name"]
            assert res["tie_count"] == 1, c["name"]


def test_picker_vectors(golden):
    # maxscore/picker_test.go: order by score desc; ties compared as SETS (tieBreakCandidates)
    for c in golden["picker"]["cases"]:
        n = len(c["scores"])
        snap = o.SnapshotData(kv_usage=np.zeros(n), queue=np.zeros(n), endpoint_cols=[[s / 100.0 for s in c["scores"]]])
        res = o.schedule_one(snap, o.make_profile([(8, 100.0)]))
        order = sorted(range(n), key=lambda m: -res["weighted"][m])[: c["max_num"]]
        k = c["tie_break_candidates"]
        assert sorted(order[:k]) == sorted(c["want_order"][:k]) and order[k:] == c["want_order"][k:], c["name"]
        top = max(c["scores"])
        assert res["tie_set"] == [m for m in range(n) if c["scores"][m] == top]
        assert res["pick"] == res["tie_set"][0]
        # seeded-random tie mode: the pick stays inside the reference's tie set and is reproducible
        picks = set()
        for r in range(64):
            pr = o.make_profile([(8, 100.0)], tie_mode=o.TIE_SEEDED_RANDOM, tie_seed=7)
            rr = o.schedule_one(snap, pr, request_index=r)
            assert rr["pick"] in res["tie_set"]
            picks.add(rr["pick"])
        assert picks == set(res["tie_set"])  # every tie member is reachable


# ---------------------------------------------------------------- prefix producer + indexer
def test_prepare_empty_index(golden):
    c = golden["prepare_empty_index"]
    idx = o.Index()
    h = o.hash_prompt(c["prompt"].encode(), o.model_seed(c["model"]), c["block_chars"], c["max_blocks"])
    assert len(h) == c["want_total"]
    assert list(idx.match(h, c["n_endpoints"])) == c["want_match"]


def test_pre_request(golden):
    c = golden["pre_request"]
    idx = o.Index()
    h = o.hash_prompt(c["prompt"].encode(), o.model_seed(c["model"]), c["block_chars"], c["max_blocks"])
    idx.add(h, c["pick"])
    for x in h:
        assert c["pick"] in idx.get(x)


def test_prefix_completion(golden):
    c = golden["prefix_completion"]
    idx = o.Index()
    seed = o.model_seed(c["model"])
    h1 = o.hash_prompt(c["first_prompt"].encode(), seed, c["block_chars"], c["max_blocks"])
    assert len(h1) == c["first_total"]
    for s in c["commit_to"]:                                      # primary pick + "prefill" profile pick (plugin.go:177-184)
        idx.add(h1, s)
    h2 = o.hash_prompt(c["second_prompt"].encode(), seed, c["block_chars"], c["max_blocks"])
    assert len(h2) == c["want_total"]
    assert list(idx.match(h2, c["n_endpoints"])) == c["want_match"]


def test_indexer_add_and_get(golden):
    c = golden["indexer"]["add_and_get"]
    idx = o.Index(c["default_lru"])
    for st in c["steps"]:
        idx.add(st["add"], 0, c["gpu_blocks"])
        assert idx.lru_len(0) == st["want_len"]
        for h, want in st.get("want_get", {}).items():
            assert sorted(idx.get(int(h))) == want


def test_indexer_remove_pod_and_eviction(golden):
    n = golden["indexer"]["remove_pod_and_eviction"]["indexer_size"]
    idx = o.Index(n)
    for j in range(n):
        idx.add([j], 1)
        idx.add([j], 2)
    assert idx.lru_len(1) == n and idx.lru_len(2) == n
    for j in range(n):
        assert idx.get(j) == {1, 2}
    idx.add([n], 1)                                               # evicts hash 0 from server1
    assert idx.lru_len(1) == n
    assert idx.get(0) == {2}
    idx.remove_pod(2)
    assert idx.get(0) == set()
    assert idx.lru_len(2) == -1 and idx.pods() == [1]
    assert idx.num_hashes() == n                                  # hashes 1..n, all only on server1
    for j in range(1, n + 1):
        assert idx.get(j) == {1}
    assert idx.lru_keys(1) == list(range(1, n + 1))               # Keys(): oldest → newest


def test_indexer_lru_semantics():
    idx = o.Index(3)
    idx.add([1, 2, 3], 0)
    idx.add([1], 0)                                               # refresh recency, no eviction
    assert idx.lru_keys(0) == [2, 3, 1]
    idx.add([4], 0)                                               # evicts 2 (oldest)
    assert idx.lru_keys(0) == [3, 1, 4] and idx.get(2) == set()
    # one Add longer than the capacity: LRU adds (with evictions) happen BEFORE the hashToPods
    # update (indexer.go:70-82) ⇒ evicted-in-this-call hashes are re-inserted into hashToPods (stale)
    idx2 = o.Index(2)
    idx2.add([10, 11, 12], 0)
    assert idx2.lru_keys(0) == [11, 12]
    assert idx2.get(10) == {0} and idx2.get(11) == {0} and idx2.get(12) == {0}
    # matchLongestPrefix stops at the first globally-unknown hash, not per pod (plugin.go:224-233)
    idx3 = o.Index(100)
    idx3.add([1, 2, 3], 0)
    idx3.add([1, 3], 1)
    assert list(idx3.match([1, 2, 3, 4, 5], 3)) == [3, 2, 0]     # pod1 counts hash 3 although it lacks hash 2
    assert list(idx3.match([9, 1, 2], 3)) == [0, 0, 0]
    # Get does not touch recency (indexer.go:86-102)
    idx4 = o.Index(2)
    idx4.add([1, 2], 0)
    idx4.get(1)
    idx4.add([3], 0)
    assert idx4.lru_keys(0) == [2, 3]


def test_batch_matches_single_and_threads_agree():
    from tests.helpers import synth_prompts, synth_snapshot, zipf_adapters
    M, R = 96, 300
    sd = synth_snapshot(M, seed=3, tie_heavy=True)
    snap = o.SnapshotData(**sd)
    prof = o.make_profile(kinds([("queue", 2), ("kv", 2), ("prefix", 3), ("lora", 1)]))
    prompts, off, _ = synth_prompts(R, prompt_len=512, groups=7, shared=256, seed=3)
    seeds = np.full(R, o.model_seed("m"), dtype=np.uint64)
    idx = o.Index(1000)
    warm = o.schedule_batch(snap, prof, idx, R, prompt_bytes=prompts, prompt_off=off, model_seed=seeds,
                            adapter_id=zipf_adapters(R, seed=3), want_hashes=True, max_blocks=16)
    idx.commit(warm["pick"], warm["hashes_out"], warm["total_blocks"])
    a = o.schedule_batch(snap, prof, idx, R, prompt_bytes=prompts, prompt_off=off, model_seed=seeds,
                         adapter_id=zipf_adapters(R, seed=3), want_match=True, want_hashes=True, max_blocks=16,
                         n_threads=1)
    b = o.schedule_batch(snap, prof, idx, R, prompt_bytes=prompts, prompt_off=off, model_seed=seeds,
                         adapter_id=zipf_adapters(R, seed=3), want_match=True, want_hashes=True, max_blocks=16,
                         n_threads=4)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert a["match_blocks"].max() > 0 and (a["total_blocks"] == 8).all()
    ad = zipf_adapters(R, seed=3)
    for r in (0, 1, 17, R - 1):
        one = o.schedule_one(snap, prof, request_index=r, adapter_id=int(ad[r]), match=a["match_blocks"][r],
                             total=int(a["total_blocks"][r]))
        assert one["pick"] == a["pick"][r] and one["score"] == a["pick_score"][r]
        assert one["tie_count"] == a["tie_count"][r]


# ---------------------------------------------------------------- token-load scorer, latency fold-in
def test_token_load_scorer(golden):
    g = golden["token_load_scorer"]
    snap = o.SnapshotData([0.0] * 3, [0] * 3, inflight_tokens=g["tokens"])
    got = o.score_single("token_load", snap, threshold=g["threshold"])
    assert np.allclose(got, g["want"], atol=g["tolerance"])
    # attribute absent => tokenLoad 0 => 1.0 (token_load.go:89-100); threshold <= 0 => default 4194304 (:57-61)
    assert np.array_equal(o.score_single("token_load", o.SnapshotData([0.0] * 2, [0] * 2), threshold=10.0), [1.0, 1.0])
    snap2 = o.SnapshotData([0.0] * 3, [0] * 3, inflight_tokens=[-5, 2097152, 1 << 40])
    assert np.array_equal(o.score_single("token_load", snap2, threshold=0.0), [1.0, 0.5, 0.0])


def test_latency_validate(golden):
    for c in golden["latency_validate"]["cases"]:
        lp = o.make_latency_params(streaming_mode=c["streaming"])
        neutral = bool(c.get("neutralize", 0)) or not c["streaming"]
        got = o.latency_validate(lp, c["ttft"], c["tpot"], c["ttft_slo"], c["tpot_slo"], c["pod_min"], neutral)
        assert got["ttft_ok"] == c["want_ttft_ok"], c["name"]
        assert got["tpot_ok"] == c["want_tpot_ok"], c["name"]
        assert got["valid"] == c["want_valid"], c["name"]
        if c.get("want_ttft_headroom_pos"):
            assert got["ttft_headroom"] > 0, c["name"]
        if "want_headroom_pos" in c:
            assert (got["headroom"] > 0) == c["want_headroom_pos"], c["name"]
        if "want_headroom" in c:
            assert got["headroom"] == c["want_headroom"], c["name"]


def test_latency_scorer(golden):
    lp = o.make_latency_params()
    for c in golden["latency_scorer"]["cases"]:
        eps = c["endpoints"]
        snap = o.SnapshotData([e[0] for e in eps], [e[1] for e in eps], [e[2] for e in eps])
        if c["info"] is None:
            got = o.score_latency_info(lp, snap, None, None, None)
        else:
            got = o.score_latency_info(lp, snap, [1] * len(eps), [i[0] for i in c["info"]], [i[1] for i in c["info"]],
                                       [i[2] for i in c["info"]])
        a = c["assert"]
        for i in a.get("nonzero", []):
            assert got[i] != 0, c["name"]
        for i in a.get("zero", []):
            assert got[i] == 0, c["name"]
        for hi, lo in a.get("greater", []):
            assert got[hi] > got[lo], c["name"]
        assert np.allclose(got, c["derived"], atol=1e-12), (c["name"], got)


def test_latency_predict_and_pipeline():
    # prediction.go:164-194, left-to-right float64; exact against the same expression in Python floats
    lp = o.make_latency_params(ttft_intercept=12.5, ttft_kv=80.0, ttft_input=0.031, ttft_waiting=7.25, ttft_running=1.5,
                               ttft_prefix=-40.0, tpot_intercept=9.0, tpot_kv=11.0, tpot_input=0.0007,
                               tpot_waiting=0.9, tpot_running=0.35, tpot_generated=0.01, streaming_mode=1)
    rng = np.random.default_rng(5)
    for _ in range(200):
        kv, inp, w, r, pf = rng.random(), int(rng.integers(0, 5000)), int(rng.integers(0, 50)), int(rng.integers(0, 90)), rng.random()
        t, p = o.latency_predict(lp, kv, inp, w, r, pf)
        assert t == ((((12.5 + 80.0 * kv) + 0.031 * float(inp)) + 7.25 * float(w)) + 1.5 * float(r)) + -40.0 * pf
        assert p == ((((9.0 + 11.0 * kv) + 0.0007 * float(inp)) + 0.9 * float(w)) + 0.35 * float(r)) + 0.01 * 1.0
    # producer + scorer == scorer over the producer's headrooms, candidates only
    M = 37
    kvs, q, run = rng.random(M), rng.integers(0, 20, M), rng.integers(0, 30, M)
    snap = o.SnapshotData(kvs, q, run, min_tpot_slo=rng.choice([0.0, 25.0, 60.0], M), dispatched=rng.integers(0, 3, M),
                          prefill_role=(rng.random(M) < 0.2).astype(np.uint8))
    match = rng.integers(0, 9, M).astype(np.uint16)
    mask = mask_from_list(M, [m for m in range(M) if m % 3])
    for ttft_slo, tpot_slo in [(0.0, 0.0), (150.0, 30.0), (400.0, 50.0), (1e6, 1e6)]:
        got, pred = o.score_latency(lp, snap, 700, ttft_slo, tpot_slo, mask, match, 8)
        th, ph, have = np.zeros(M), np.zeros(M), np.ones(M, np.uint8)
        for m in range(M):
            t, p = o.latency_predict(lp, kvs[m], 700, int(q[m]), int(run[m]), match[m] / 8)
            assert (t, p) == tuple(pred[m])
            v = o.latency_validate(lp, t, p, ttft_slo, tpot_slo, snap.min_tpot_slo[m], bool(snap.prefill_role[m]))
            th[m], ph[m] = v["ttft_headroom"], v["headroom"]
        want = o.score_latency_info(lp, snap, have, th, ph, snap.dispatched, mask)
        assert np.array_equal(got, want, equal_nan=True)
        cand = ~np.isnan(got)
        assert cand.sum() == sum(1 for m in range(M) if m % 3) and (got[cand] >= 0).all() and (got[cand] <= 1.01).all()


# ---------------------------------------------------------------- stochastic pickers: distribution level
def _scores_snapshot(scores):
    # scores arrive through a custom per-endpoint column scorer with weight 100: score = clamp(col) * 100
    M = len(scores)
    return o.SnapshotData([0.0] * M, [0] * M, endpoint_cols=[np.array(scores, np.float64) / 100.0])


def test_stochastic_pickers_distribution(golden):
    g = golden["stochastic_pickers"]
    n, tol = g["iterations"], g["tolerance"]
    for c in g["weighted"]:
        sc = np.array(c["scores"], np.float64)
        prof = o.make_profile([(o.SCORER_ENDPOINT_COL0, 100.0)], tie_seed=1234, pick_mode=o.PICK_WEIGHTED_RANDOM)
        res = o.schedule_batch(_scores_snapshot(sc), prof, None, n)
        freq = np.bincount(res["pick"], minlength=len(sc)) / n
        assert np.abs(freq - sc / sc.sum()).max() <= tol, (c["name"], freq)
        assert (freq[sc == 0] == 0).all(), c["name"]                       # key 0: never selected (picker.go:127-131)
        assert (res["tie_count"] == (sc > 0).sum()).all()
        assert np.array_equal(res["pick_score"], sc[res["pick"]])
    for c in g["random"]:
        sc = np.array(c["scores"], np.float64)
        prof = o.make_profile([(o.SCORER_ENDPOINT_COL0, 100.0)], tie_seed=99, pick_mode=o.PICK_RANDOM)
        res = o.schedule_batch(_scores_snapshot(sc), prof, None, n)
        freq = np.bincount(res["pick"], minlength=len(sc)) / n
        assert np.abs(freq - 1.0 / len(sc)).max() <= tol, (c["name"], freq)
    # all scores zero: the weighted-random picker delegates to the random picker (picker.go:113-116)
    prof = o.make_profile([(o.SCORER_ENDPOINT_COL0, 100.0)], tie_seed=7, pick_mode=o.PICK_WEIGHTED_RANDOM)
    res = o.schedule_batch(_scores_snapshot([0, 0, 0, 0]), prof, None, n)
    freq = np.bincount(res["pick"], minlength=4) / n
    assert np.abs(freq - 0.25).max() <= tol and (res["tie_count"] == 4).all()


def test_stochastic_generator_properties():
    L = o.lib()
    import math
    for u in (1.0, 0.5, 0.7, 1e-10, 2.0 ** -53, 0.9999999, 1.0 / 3.0, 0.70710678118654757, 0.7071067811865476):
        assert abs(L.orc_neg_log(u) + math.log(u)) <= 4e-16 * max(1.0, abs(math.log(u))), u
    us = np.array([L.orc_uniform01(42, r, r % 7) for r in range(50000)])
    assert us.min() > 0.0 and us.max() <= 1.0
    assert abs(us.mean() - 0.5) < 0.01 and abs(us.var() - 1.0 / 12.0) < 0.005
    hist = np.histogram(us, bins=20, range=(0, 1))[0] / len(us)
    assert np.abs(hist - 0.05).max() < 0.006


# ---------------------------------------------------------------- len(strings.Fields(prompt)), the latency path's token count
FIELDS_CASES = [
    (b"", 0), (b"   ", 0), (b"a", 1), (b"  a  b\tc\n", 3), (b"a\x0bb\x0cc\rd", 4),
    (b"one\xc2\xa0two\xe2\x80\x83three\xe3\x80\x80four", 4),      # NBSP, EM SPACE, IDEOGRAPHIC SPACE
    (b"x\xe1\x9a\x80y\xe2\x80\xa8z\xe2\x80\xafw\xe2\x81\x9f\xe2\x80\xa9q", 5),  # OGHAM, LINE SEP, NNBSP, MMSP, PARA SEP
    (b"\xe2\x80\x8b", 1),                                        # ZERO WIDTH SPACE is not White_Space
    (b"a\xc2\x85b", 2),                                          # NEL
    (b"\xf0\xe2\x80\x80x", 2),                                   # invalid lead consumes ONE byte, then EN QUAD
    (b"a\xe2\x80", 1), (b"\xc2\x85\xc2", 1), (b"\xe2\x80\x80", 0), (b"\x1c\x1d\x1e\x1f", 1),  # FS..US are not spaces in Go
]


def _fields_by_byte_patterns(s: bytes) -> int:
    """The rule the device kernel and the C++ host mirror use: a byte is space iff ASCII space or inside one of the
    UTF-8 encodings of the White_Space runes; no rune decoding."""
    n = len(s)
    sp = [False] * n
    for i, b in enumerate(s):
        if b in (9, 10, 11, 12, 13, 32):
            sp[i] = True
        if b == 0xC2 and i + 1 < n and s[i + 1] in (0x85, 0xA0):
            sp[i] = sp[i + 1] = True
        if i + 2 < n:
            b1, b2 = s[i + 1], s[i + 2]
            if ((b == 0xE1 and b1 == 0x9A and b2 == 0x80) or
                    (b == 0xE2 and b1 == 0x80 and (0x80 <= b2 <= 0x8A or b2 in (0xA8, 0xA9, 0xAF))) or
                    (b == 0xE2 and b1 == 0x81 and b2 == 0x9F) or (b == 0xE3 and b1 == 0x80 and b2 == 0x80)):
                sp[i] = sp[i + 1] = sp[i + 2] = True
    return sum(1 for i in range(n) if not sp[i] and (i == 0 or sp[i - 1]))


def test_count_fields_go_semantics():
    for s, want in FIELDS_CASES:
        assert o.count_fields(s) == want, s
        assert _fields_by_byte_patterns(s) == want, s
    # the byte-pattern rule == Go's rune decoding, on adversarial byte soup
    import random
    rnd = random.Random(7)
    alphabet = [0x20, 0x09, 0x41, 0x42, 0xC2, 0x85, 0xA0, 0xE1, 0x9A, 0x80, 0xE2, 0x81, 0x9F, 0xE3, 0x8A, 0xA8, 0xA9, 0xAF,
                0xF0, 0xF4, 0xED, 0xE0, 0xC0, 0xFF, 0x0A, 0x8B, 0x90, 0xBF]
    for _ in range(20000):
        s = bytes(rnd.choice(alphabet) for _ in range(rnd.randint(0, 24)))
        assert o.count_fields(s) == _fields_by_byte_patterns(s), s


# ---------------------------------------------------------------- the "Go-shape" CPU baseline computes the same results
def test_goshape_baseline_matches_oracle():
    from tests.helpers import synth_prompts, synth_snapshot, zipf_adapters
    M, R = 96, 300
    sd = synth_snapshot(M, seed=2, tie_heavy=True)
    snap = o.SnapshotData(**sd)
    prof = o.make_profile(kinds([("queue", 2), ("kv", 2), ("prefix", 3), ("lora", 1)]))
    prompts, off, _ = synth_prompts(R, prompt_len=512, groups=7, shared=256, seed=2)
    seeds = np.full(R, o.model_seed("m"), np.uint64)
    ad = zipf_adapters(R, seed=2)
    idx = o.Index()
    warm = o.schedule_batch(snap, prof, idx, R, prompt_bytes=prompts, prompt_off=off, model_seed=seeds, adapter_id=ad,
                            want_hashes=True)
    idx.commit(warm["pick"], warm["hashes_out"], warm["total_blocks"])
    kw = dict(prompt_bytes=prompts, prompt_off=off, model_seed=seeds, adapter_id=ad)
    want = o.schedule_batch(snap, prof, idx, R, want_tie_set=True, **kw)
    for threads, seed in ((1, 0), (3, 99)):
        got = o.schedule_batch(snap, prof, idx, R, goshape=True, n_threads=threads, shuffle_seed=seed, **kw)
        assert np.array_equal(got["pick_score"].view(np.uint64), want["pick_score"].view(np.uint64))
        assert np.array_equal(got["tie_count"], want["tie_count"])
        ts = want["tie_set"]
        assert ((ts[np.arange(R), got["pick"] >> 5] >> (got["pick"] & 31).astype(np.uint32)) & 1).all()
    # a tie-heavy profile: the shuffle picks some member of the arg-max set, the set's size is the same
    sd["kv_usage"] = np.round(sd["kv_usage"], 1)
    snap2 = o.SnapshotData(**sd)
    prof2 = o.make_profile(kinds([("kv", 1), ("lora", 1)]))
    want2 = o.schedule_batch(snap2, prof2, None, R, adapter_id=ad, want_tie_set=True)
    got2 = o.schedule_batch(snap2, prof2, None, R, adapter_id=ad, prompt_bytes=prompts, prompt_off=off, goshape=True, shuffle_seed=5)
    assert np.array_equal(got2["tie_count"], want2["tie_count"]) and want2["tie_count"].max() > 1
    ts2 = want2["tie_set"]
    assert ((ts2[np.arange(R), got2["pick"] >> 5] >> (got2["pick"] & 31).astype(np.uint32)) & 1).all()
    assert (got2["pick"] != want2["pick"]).any()  # the shuffle does not always land on the lowest index of a tie set


# ---------------------------------------------------------------- device-side filters of the latency profile
def filter_case_inputs(kind, c):
    """Encodes a reference filter test as engine/oracle inputs.  Headrooms / TTFTs ride on the identity model
    TTFT = WaitingQueueSize, TPOT = RunningRequestsSize with both SLOs = 1000; prefix matches of 100 blocks come from an
    index in which endpoint m holds the first `match` block hashes of the request."""
    eps = c["endpoints"]
    M = len(eps)
    hashes = (np.arange(100, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
    if kind == "tier":
        has_pred = all(e[2] for e in eps)
        queue = [1000 - int(e[0]) for e in eps]
        running = [1000 - int(e[1]) for e in eps]
        adds = []
        filters = [(o.FILTER_SLO_HEADROOM_TIER, (c["epsilon"],))]
    else:
        has_pred = True
        queue = [int(e[1]) for e in eps]
        running = [0] * M
        adds = [(m, int(e[0])) for m, e in enumerate(eps) if e[0] > 0]
        filters = [(o.FILTER_PREFIX_AFFINITY, tuple(c["params"]))]
    lat = dict(ttft_waiting=1.0, tpot_running=1.0, streaming_mode=1, has_predictions=1 if has_pred else 0)
    return dict(M=M, queue=queue, running=running, adds=adds, filters=filters, lat=lat, hashes=hashes)


def test_filters_reference_cases(golden):
    for kind, key in (("tier", "slo_headroom_tier"), ("affinity", "prefix_cache_affinity")):
        for c in golden["filters"][key]["cases"]:
            inp = filter_case_inputs(kind, c)
            M = inp["M"]
            snap = o.SnapshotData(np.zeros(M), inp["queue"], inp["running"])
            idx = o.Index()
            for m, n in inp["adds"]:
                idx.add(inp["hashes"][:n], m)
            prof = o.make_profile([(o.SCORER_LATENCY, 1.0)], latency=o.make_latency_params(**inp["lat"]), filters=inp["filters"], tie_seed=3)
            res = o.schedule_batch(snap, prof, idx, 1, hashes_in=inp["hashes"][None, :], n_hashes_in=np.array([100], np.uint16),
                                   ttft_slo=[1000.0], tpot_slo=[1000.0], want_filter_mask=True, want_match=True)
            kept = [m for m in range(M) if (res["filter_mask_out"][0][m >> 5] >> (m & 31)) & 1]
            assert kept == c["want"], (c["name"], kept)
            assert res["pick"][0] in kept


def test_filter_draws_are_bernoulli():
    # both tiers present, epsilon 0.3: the negative tier is selected ~30 % of the time
    lp = o.make_latency_params(ttft_waiting=1.0, tpot_running=1.0, streaming_mode=1)
    snap = o.SnapshotData(np.zeros(3), [900, 800, 1100], [950, 920, 1050])
    prof = o.make_profile([(o.SCORER_LATENCY, 1.0)], latency=lp, filters=[(o.FILTER_SLO_HEADROOM_TIER, (0.3,))], tie_seed=11)
    R = 20000
    res = o.schedule_batch(snap, prof, None, R, ttft_slo=np.full(R, 1000.0), tpot_slo=np.full(R, 1000.0), want_filter_mask=True)
    neg = (res["filter_mask_out"][:, 0] == 0b100)
    pos = (res["filter_mask_out"][:, 0] == 0b011)
    assert (neg | pos).all() and abs(neg.mean() - 0.3) < 0.02
    assert (res["pick"][neg] == 2).all() and (res["pick"][pos] != 2).all()


# ---------------------------------------------------------------- the latency kernel's reciprocal fast path
def test_latency_fast_path_never_changes_the_quantised_weight():
    """score_matrix.cu normalises with d * rcp(range) instead of d / range and redoes a pair exactly only when
    v = (1 - combined) * 100 (or combined * 100) lies within 1e-9 of an integer.  Claim: outside that band the truncated
    weight int(v) is identical.  Checked here in IEEE float64 (numpy: every operation individually rounded, like the _rn
    intrinsics) on adversarial inputs — ranges over 15 decades, weights on and off the simplex, headrooms at the extremes."""
    rng = np.random.Generator(np.random.PCG64(2024))
    n = 2_000_000
    worst = 0.0
    for most in (False, True):
        rg_t = 10.0 ** rng.uniform(-8.9, 6, n)
        rg_p = 10.0 ** rng.uniform(-8.9, 6, n)
        # numerators d = |h| - min in [0, range], with mass at the ends and on k/100 grid points
        frac_t = np.where(rng.random(n) < 0.3, rng.integers(0, 101, n) / 100.0, rng.random(n))
        frac_p = np.where(rng.random(n) < 0.3, rng.integers(0, 101, n) / 100.0, rng.random(n))
        d_t, d_p = np.minimum(frac_t * rg_t, rg_t), np.minimum(frac_p * rg_p, rg_p)
        alpha = np.where(rng.random(n) < 0.5, 0.8, rng.random(n))
        beta = np.where(rng.random(n) < 0.5, 1.0 - alpha, rng.random(n))
        exact_c = alpha * (d_t / rg_t) + beta * (d_p / rg_p)
        fast_c = alpha * (d_t * (1.0 / rg_t)) + beta * (d_p * (1.0 / rg_p))
        v_exact = exact_c * 100.0 if most else (1.0 - exact_c) * 100.0
        v_fast = fast_c * 100.0 if most else (1.0 - fast_c) * 100.0
        worst = max(worst, float(np.abs(v_exact - v_fast).max()))
        decided = np.abs(v_fast - np.rint(v_fast)) >= 1e-9          # the kernel keeps the fast result only here
        assert np.array_equal(np.trunc(v_fast[decided]), np.trunc(v_exact[decided]))
        assert decided.mean() > 0.5                                  # and the exact redo stays the rare path
    assert worst < 2e-13, worst


def test_composite_fast_path_never_changes_the_rounded_weight():
    """Same argument for the composite fallback (plugin.go:346-360): relQueue via d * rcp(maxQ) instead of d / maxQ, exact
    redo only when 100 * composite is within 1e-9 of a half-integer (math.Round's decision points) or beyond 1e4."""
    rng = np.random.Generator(np.random.PCG64(7))
    n = 2_000_000
    maxq = rng.integers(1, 100000, n).astype(np.float64)
    dq = np.floor(rng.random(n) * (maxq + 1)).clip(0, maxq)
    w = rng.random((3, n))
    w /= w.sum(axis=0)                      # the normalised weights (each in [0,1])
    kv_free = np.where(rng.random(n) < 0.2, rng.integers(0, 101, n) / 100.0, rng.random(n))
    prefix = rng.integers(0, 33, n) / 32.0
    ck, pt = w[0] * kv_free, w[2] * prefix
    v_exact = 100.0 * ((ck + w[1] * (dq / maxq)) + pt)
    v_fast = 100.0 * ((ck + w[1] * (dq * (1.0 / maxq))) + pt)
    assert np.abs(v_exact - v_fast).max() < 1e-12
    decided = np.abs((v_fast - np.floor(v_fast)) - 0.5) >= 1e-9
    half_away = lambda x: np.floor(x + 0.5)  # == math.Round for x >= 0
    assert np.array_equal(half_away(v_fast[decided]), half_away(v_exact[decided]))
    assert decided.mean() > 0.99


# ---------------------------------------------------------------- the identity behind the sparse pick kernel (pick_sparse.cu)
def test_sparse_pick_identity_holds_on_the_oracle():
    """pick_sparse.cu never scores the endpoints without a prefix match: it combines the per-adapter summary of the
    zero-match score map G[a][m] with the few "exception" endpoints.  The claim (top of pick_sparse.cu), checked here purely
    on the CPU oracle's float64 score maps for tie-heavy snapshots:
        max_m S = max(gmax, T),  T = max over exceptions of S
        T > gmax : arg-max set = exceptions attaining T
        T < gmax : arg-max set = precomputed set (no exception belongs to it)
        T == gmax: (precomputed set minus its exception members) + exceptions attaining gmax"""
    from tests.helpers import synth_snapshot, zipf_adapters
    rng = np.random.Generator(np.random.PCG64(99))
    M, R = 160, 600
    for scorers in ([("queue", 2), ("kv", 2), ("prefix", 3), ("lora", 1)], [("kv", 1), ("prefix", 0.5)], [("lora", 1), ("prefix", 2), ("kv", 1)]):
        sd = synth_snapshot(M, seed=int(rng.integers(0, 1000)), tie_heavy=True)
        sd["kv_usage"] = np.round(sd["kv_usage"], 1)
        snap = o.SnapshotData(**sd)
        prof = o.make_profile(kinds(scorers))
        ad = zipf_adapters(R, seed=1)
        total = 8
        for r in range(R):
            match = np.zeros(M, np.uint16)
            nexc = int(rng.integers(0, 12))
            exc = rng.choice(M, nexc, replace=False)
            match[exc] = rng.integers(1, total + 1, nexc)
            full = o.schedule_one(snap, prof, adapter_id=int(ad[r]), match=match, total=total)
            zero = o.schedule_one(snap, prof, adapter_id=int(ad[r]), match=np.zeros(M, np.uint16), total=total)
            S, G = full["weighted"], zero["weighted"]
            assert (S[exc] >= G[exc]).all()                    # monotonicity: a match can only raise a score (weight >= 0)
            gmax, gset = G.max(), set(np.nonzero(G == G.max())[0])
            T = S[exc].max() if nexc else -np.inf
            assert full["score"] == max(gmax, T)
            eset = set(int(e) for e in exc)
            if T > gmax:
                want = {int(e) for e in exc if S[e] == T}
            elif T < gmax:
                assert not (gset & eset) or all(S[e] < gmax or G[e] < gmax for e in gset & eset)
                want = {m for m in gset if m not in eset} | {int(e) for e in exc if S[e] == gmax}
            else:
                want = {m for m in gset if m not in eset} | {int(e) for e in exc if S[e] == gmax}
            assert set(full["tie_set"]) == want, (r, scorers)


# ---------------------------------------------------------------- a second, independent restatement (NumPy) pins the C oracle
def test_numpy_restatement_agrees_with_c_oracle():
    """The Go reference cannot run here, so the C oracle is additionally cross-checked against an independent NumPy
    restatement of the same reference lines (queue.go:78-108, kvcache_utilization.go:76-82, lora_affinity.go:76-102,
    prefix/plugin.go:95-117, scheduler_profile.go:151-202, maxscore/picker.go:87-115) on random snapshots and masks."""
    from tests.helpers import synth_snapshot, zipf_adapters
    rng = np.random.Generator(np.random.PCG64(314))
    for trial in range(6):
        M = int(rng.integers(3, 200))
        sd = synth_snapshot(M, seed=trial, tie_heavy=trial % 2 == 0)
        snap = o.SnapshotData(**sd)
        order = [("queue", 2.0), ("kv", 2.0), ("prefix", 3.0), ("lora", 1.0), ("running", 0.5)]
        rng.shuffle(order)
        weights = [(k, float(w) * float(rng.choice([1.0, -0.5, 1.5]))) for k, w in order]
        prof = o.make_profile(kinds(weights))
        for r in range(40):
            cand = rng.random(M) < rng.choice([0.2, 0.7, 1.0])
            if not cand.any():
                cand[int(rng.integers(0, M))] = True
            total = int(rng.integers(0, 9))
            match = rng.integers(0, total + 1, M).astype(np.uint16)
            ad = int(zipf_adapters(1, seed=r)[0])
            got = o.schedule_one(snap, prof, adapter_id=ad, mask=mask_from_list(M, list(np.nonzero(cand)[0])), match=match, total=total)
            acc = np.zeros(M)
            for k, w in weights:
                if k in ("queue", "running"):
                    q = np.asarray(sd["queue" if k == "queue" else "running"], np.int64)[cand]
                    sc = np.ones(M)
                    if q.max() != q.min():
                        full = np.asarray(sd["queue" if k == "queue" else "running"], np.int64)
                        sc = (q.max() - full).astype(np.float64) / np.float64(q.max() - q.min())
                elif k == "kv":
                    sc = 1.0 - np.asarray(sd["kv_usage"])
                elif k == "prefix":
                    sc = match.astype(np.float64) / np.float64(total) if total else np.zeros(M)
                else:
                    act = wai = np.zeros(M, np.uint64)  # a target model outside the adapter dictionary is on no endpoint
                    if ad >= 0:
                        w_, b_ = ad >> 6, np.uint64(ad & 63)
                        act = (sd["lora_active"][:, w_] >> b_) & np.uint64(1)
                        wai = (sd["lora_waiting"][:, w_] >> b_) & np.uint64(1)
                    sc = np.where(act == 1, 1.0, np.where(sd["lora_nmodels"] < sd["lora_max"], 0.8, np.where(wai == 1, 0.6, 0.0)))
                acc = acc + np.clip(sc, 0.0, 1.0) * np.float64(w)
            best = acc[cand].max()
            ties = [int(m) for m in np.nonzero(cand & (acc == best))[0]]
            assert got["score"] == best and got["tie_set"] == ties and got["pick"] == ties[0], (trial, r)
            assert np.array_equal(got["weighted"][cand].view(np.uint64), acc[cand].view(np.uint64))


def _latency_scorer_py(th, ph, disp, cand, ttft_w=0.8, tpot_w=0.2, most=False):
    """scorer/latency/plugin.go:144-318 restated with Python lists (independent of oracle.c)."""
    import math
    M = len(th)
    scores = [0.0 if cand[m] else float("nan") for m in range(M)]
    data = [m for m in range(M) if cand[m]]

    def bucket(ms, force_least):
        s = ttft_w + tpot_w
        a, b = (1.0, 0.0) if s <= 0 else (ttft_w / s, tpot_w / s)
        at, ap = [abs(th[m]) for m in ms], [abs(ph[m]) for m in ms]
        rt, rp = max(at) - min(at), max(ap) - min(ap)
        if rt <= 1e-9 and rp > 1e-9:
            a, b = 0.0, 1.0
        elif rp <= 1e-9 and rt > 1e-9:
            a, b = 1.0, 0.0
        for m, x, y in zip(ms, at, ap):
            nt = (x - min(at)) / rt if rt > 1e-9 else 0.5
            np_ = (y - min(ap)) / rp if rp > 1e-9 else 0.5
            c = a * nt + b * np_
            w = int(c * 100.0) + 1 if (most and not force_least) else int((1.0 - c) * 100.0) + 1
            scores[m] = float(w) / 100.0

    positive = [m for m in data if not (th[m] < 0 or ph[m] < 0)]
    negative = [m for m in data if th[m] < 0 or ph[m] < 0]
    if positive:
        bucket(positive, False)
        return scores
    idle = [m for m in negative if disp[m] == 0]
    if idle:
        bucket(idle, True)
        return scores
    for sel in (lambda m: not th[m] < 0 and ph[m] < 0, lambda m: th[m] < 0 and not ph[m] < 0, lambda m: th[m] < 0 and ph[m] < 0):
        b_ = [m for m in negative if sel(m)]
        if b_:
            bucket(b_, True)
            return scores
    return scores


def test_python_restatement_of_latency_scorer_agrees_with_c_oracle():
    rng = np.random.Generator(np.random.PCG64(2718))
    for trial in range(400):
        M = int(rng.integers(1, 40))
        scale = rng.choice([1.0, 50.0, 1e-10])            # incl. ranges below the scorer's eps
        th = np.round(rng.normal(rng.choice([-30, 0, 40]), 25, M), int(rng.integers(0, 3))) * scale
        ph = np.round(rng.normal(rng.choice([-5, 0, 8]), 6, M), int(rng.integers(0, 3))) * scale
        if trial % 7 == 0:
            ph[:] = 0.0                                   # neutralised TPOT (non-streaming)
        disp = rng.integers(0, 3, M).astype(np.int32) if trial % 3 else np.ones(M, np.int32)
        cand = rng.random(M) < rng.choice([0.3, 1.0])
        if not cand.any():
            cand[0] = True
        most = bool(trial % 2)
        lp = o.make_latency_params(strategy_most=1 if most else 0, ttft_weight=float(rng.choice([0.8, 0.0, 1.0])),
                                   tpot_weight=float(rng.choice([0.2, 0.0, 3.0])))
        snap = o.SnapshotData(np.zeros(M), np.zeros(M, np.int64))
        got = o.score_latency_info(lp, snap, np.ones(M, np.uint8), th, ph, disp, mask=mask_from_list(M, list(np.nonzero(cand)[0])))
        want = _latency_scorer_py(list(th), list(ph), list(disp), list(cand), lp.ttft_weight, lp.tpot_weight, most)
        assert np.array_equal(np.array(want), got, equal_nan=True), (trial, want, got)


def test_python_restatement_of_filters_agrees_with_c_oracle():
    """prefixcacheaffinity/plugin.go:105-151 and sloheadroomtier/plugin.go:82-137 restated with Python lists, chained in random
    orders with random parameters; the draws come from the shared counter-based generator (orc_uniform01)."""
    L = o.lib()
    rng = np.random.Generator(np.random.PCG64(1618))
    lp = o.make_latency_params(ttft_waiting=1.0, tpot_running=1.0, streaming_mode=1)   # TTFT = waiting, TPOT = running
    for trial in range(300):
        M = int(rng.integers(1, 30))
        total = int(rng.integers(0, 11))
        match = rng.integers(0, total + 1, M).astype(np.uint16)
        queue = rng.integers(0, 400, M).astype(np.int64)
        running = rng.integers(0, 60, M).astype(np.int64)
        ttft_slo, tpot_slo = float(rng.choice([0.0, 150.0, 500.0])), float(rng.choice([0.0, 30.0, 100.0]))
        cand = rng.random(M) < rng.choice([0.4, 1.0])
        filters = []
        for _ in range(int(rng.integers(1, 4))):
            if rng.random() < 0.5:
                filters.append((o.FILTER_PREFIX_AFFINITY, (float(rng.choice([0.0, 0.3, 0.8, 1.0])), float(rng.choice([0.0, 0.3, 1.0])),
                                                           float(rng.choice([0.0, 20.0, 5000.0])))))
            else:
                filters.append((o.FILTER_SLO_HEADROOM_TIER, (float(rng.choice([0.0, 0.3, 1.0])),)))
        seed, req = int(rng.integers(0, 2 ** 40)), int(rng.integers(0, 10 ** 6))
        # ---- Python restatement ----
        cur = [m for m in range(M) if cand[m]]
        score = lambda m: (match[m] / total) if total > 0 else 0.0
        th = {m: ttft_slo - float(queue[m]) for m in range(M)}
        ph = {m: tpot_slo - float(running[m]) for m in range(M)}
        for f, (kind, par) in enumerate(filters):
            if not cur:
                break
            u = L.orc_uniform01(seed, req, -(f + 1))
            if kind == o.FILTER_PREFIX_AFFINITY:
                thr, pexp, maxpen = par
                if len(cur) <= 1 or thr <= 0 or u < pexp:
                    continue
                sticky = [m for m in cur if score(m) >= thr]
                non = [m for m in cur if score(m) < thr]
                if not sticky:
                    continue
                if maxpen > 0 and non and min(float(queue[m]) for m in sticky) - min(float(queue[m]) for m in non) > maxpen:
                    continue
                cur = sticky
            else:
                if len(cur) <= 1:
                    continue
                pos = [m for m in cur if th[m] >= 0 and ph[m] >= 0]
                neg = [m for m in cur if not (th[m] >= 0 and ph[m] >= 0)]
                if pos and neg:
                    cur = neg if u < par[0] else pos
        # ---- C oracle ----
        snap = o.SnapshotData(np.zeros(M), queue, running)
        prof = o.make_profile([(o.SCORER_LATENCY, 1.0)], latency=lp, filters=filters, tie_seed=seed)
        out = np.zeros(max((M + 31) // 32, 1), np.uint32)
        lr = o.LatencyRequest(0, ttft_slo, tpot_slo)
        import ctypes as C
        L.orc_apply_filters.argtypes = [C.POINTER(o.Snapshot), C.POINTER(o.Profile), C.c_int64, C.c_void_p, C.c_void_p, C.c_int32,
                                        C.POINTER(o.LatencyRequest), C.c_void_p]
        mask = mask_from_list(M, [m for m in range(M) if cand[m]])
        L.orc_apply_filters(C.byref(snap.struct), C.byref(prof), req, mask.ctypes.data, match.ctypes.data, total, C.byref(lr),
                            out.ctypes.data)
        kept = [m for m in range(M) if (out[m >> 5] >> (m & 31)) & 1]
        assert kept == cur, (trial, filters, kept, cur)
