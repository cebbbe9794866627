"""N>1 host logic on CPU (gloo, world_size 2): one broadcast of the packed snapshot tile, request sharding with
request_base, gather of the picks — the union of the shards must equal the unsharded batch bit for bit
(the CPU oracle stands in for the per-rank engine; the GPU bench uses the same helpers with NCCL)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _pkg
from oracle import oracle_py as o
from tests.helpers import kinds, synth_prompts, synth_snapshot, zipf_adapters

SCORERS = [("queue", 2), ("kv", 2), ("prefix", 3), ("lora", 1)]
M, R_TOTAL = 160, 1001


LAT = dict(ttft_intercept=12.5, ttft_kv=80.0, ttft_input=0.031, ttft_waiting=7.25, ttft_running=1.5, ttft_prefix=-40.0,
           tpot_intercept=9.0, tpot_kv=11.0, tpot_input=0.0007, tpot_waiting=0.9, tpot_running=0.35, tpot_generated=0.01,
           streaming_mode=1)
OPT = ("min_tpot_slo", "dispatched", "prefill_role")


def _worker_latency_weighted_random(rank, world, port, q):
    """Latency-scorer profile + weighted-random picker: the per-endpoint producer state travels in the same tile, and the
    counter-based generator keyed by request_base makes the union of the shards equal the unsharded batch."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib.util
    spec = importlib.util.spec_from_file_location("sharding", os.path.join(_pkg.PKG_DIR, "sharding.py"))
    sharding = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sharding)
    layout, total = sharding.snapshot_layout(M, 1, OPT)
    if rank == 0:
        rng = np.random.Generator(np.random.PCG64(3))
        snap = synth_snapshot(M, seed=4)
        snap["min_tpot_slo"] = rng.choice([0.0, 22.0, 60.0], M)
        snap["dispatched"] = rng.integers(0, 3, M).astype(np.int32)
        snap["prefill_role"] = (rng.random(M) < 0.2).astype(np.uint8)
        buf, layout0 = sharding.pack_snapshot(snap)
        assert layout0 == layout
        t = torch.from_numpy(buf)
    else:
        t = torch.zeros(total, dtype=torch.uint8)
    dist.broadcast(t, src=0)
    snap = sharding.unpack_snapshot(t.numpy(), layout)
    rng = np.random.Generator(np.random.PCG64(8))  # the same global batch on every rank
    req = dict(input_tokens=rng.integers(0, 4000, R_TOTAL).astype(np.int32), ttft_slo=rng.choice([0.0, 150.0, 400.0], R_TOTAL),
               tpot_slo=rng.choice([0.0, 25.0, 60.0], R_TOTAL))
    lo, hi = sharding.shard_range(R_TOTAL, rank, world)
    osnap = o.SnapshotData(**snap)
    prof = o.make_profile([(o.SCORER_LATENCY, 1.0)], tie_seed=5, pick_mode=o.PICK_WEIGHTED_RANDOM,
                          latency=o.make_latency_params(**LAT),
                          filters=[(o.FILTER_SLO_HEADROOM_TIER, (0.3,))])  # the filter's draw is keyed by the global request index too
    res = o.schedule_batch(osnap, prof, None, hi - lo, request_base=lo, **{k: v[lo:hi] for k, v in req.items()})
    width = (R_TOTAL + world - 1) // world + 1
    mine = torch.full((width,), -7, dtype=torch.int32)
    mine[: hi - lo] = torch.from_numpy(res["pick"])
    allp = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allp, mine)
    if rank == 0:
        picks = np.concatenate([allp[r][: sharding.shard_range(R_TOTAL, r, world)[1] - sharding.shard_range(R_TOTAL, r, world)[0]].numpy()
                                for r in range(world)])
        whole = o.schedule_batch(osnap, prof, None, R_TOTAL, **req)
        q.put((bool(np.array_equal(picks, whole["pick"])), int(len(np.unique(whole["pick"])))))
    dist.destroy_process_group()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib.util
    spec = importlib.util.spec_from_file_location("sharding", os.path.join(_pkg.PKG_DIR, "sharding.py"))
    sharding = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sharding)

    # rank 0 owns the snapshot; ONE broadcast of the packed tile
    layout, total = sharding.snapshot_layout(M, 1)
    if rank == 0:
        snap = synth_snapshot(M, seed=9, tie_heavy=True)
        snap["kv_usage"] = np.round(snap["kv_usage"], 1)  # many exact ties: the shard-invariant tie priority matters
        buf, layout0 = sharding.pack_snapshot(snap)
        assert layout0 == layout
        t = torch.from_numpy(buf)
    else:
        t = torch.zeros(total, dtype=torch.uint8)
    dist.broadcast(t, src=0)
    snap = sharding.unpack_snapshot(t.numpy(), layout)
    # every rank generates the same global batch and takes its shard
    prompts, off, _ = synth_prompts(R_TOTAL, prompt_len=512, groups=11, shared=256, seed=9)
    ad = zipf_adapters(R_TOTAL, seed=9)
    seeds = np.full(R_TOTAL, o.model_seed("m"), np.uint64)
    lo, hi = sharding.shard_range(R_TOTAL, rank, world)
    osnap = o.SnapshotData(**snap)
    prof = o.make_profile(kinds(SCORERS), tie_mode=o.TIE_SEEDED_RANDOM, tie_seed=77)
    idx = o.Index()
    idx.add(o.hash_prompt(bytes(prompts[off[0]:off[1]]), int(seeds[0]), 64, 256), 3)  # identical replicated index
    res = o.schedule_batch(osnap, prof, idx, hi - lo, prompt_bytes=prompts[off[lo]:off[hi]], prompt_off=off[lo:hi + 1] - off[lo],
                           model_seed=seeds[lo:hi], adapter_id=ad[lo:hi], request_base=lo)
    # gather the shards' picks on every rank (variable shard sizes → pad)
    width = (R_TOTAL + world - 1) // world + 1
    mine = torch.full((width,), -7, dtype=torch.int32)
    mine[: hi - lo] = torch.from_numpy(res["pick"])
    allp = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allp, mine)
    if rank == 0:
        picks = np.concatenate([allp[r][: sharding.shard_range(R_TOTAL, r, world)[1] - sharding.shard_range(R_TOTAL, r, world)[0]].numpy()
                                for r in range(world)])
        whole = o.schedule_batch(osnap, prof, idx, R_TOTAL, prompt_bytes=prompts, prompt_off=off, model_seed=seeds, adapter_id=ad)
        q.put((bool(np.array_equal(picks, whole["pick"])), int(whole["tie_count"].max())))
    dist.destroy_process_group()


def _worker_closed_loop(rank, world, port, q):
    """The closed loop over two ranks: schedule the own shard -> gather_commit_stream (picks, counts, block hashes in rank
    order) -> replay ALL shards into the own index replica -> next batch sees the commits.  Every replica must equal the index
    of a single process that ran the unsharded batches (LRU keys of every endpoint, and therefore the next batch's picks)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib.util
    spec = importlib.util.spec_from_file_location("sharding", os.path.join(_pkg.PKG_DIR, "sharding.py"))
    sharding = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sharding)
    Mc, Rs, B, K = 48, 96, 8, 3           # endpoints, requests per rank and batch, blocks per prompt, batches
    prof = o.make_profile(kinds(SCORERS))
    seedv = o.model_seed("m")
    idx = o.Index(default_lru=40)        # small LRUs: the replay also evicts
    ref = o.Index(default_lru=40) if rank == 0 else None
    picks_seen = []
    for k in range(K):
        snap = synth_snapshot(Mc, seed=100 + k)          # a fresh snapshot per batch (every rank builds the same one)
        osnap = o.SnapshotData(**snap)
        prompts, off, _ = synth_prompts(world * Rs, prompt_len=B * 64, groups=5, shared=4 * 64, seed=200 + k, prefix_seed=3)
        ad = zipf_adapters(world * Rs, seed=300 + k)
        seeds = np.full(world * Rs, seedv, np.uint64)
        lo, hi = rank * Rs, (rank + 1) * Rs
        res = o.schedule_batch(osnap, prof, idx, Rs, prompt_bytes=prompts[off[lo]:off[hi]], prompt_off=off[lo:hi + 1] - off[lo],
                               model_seed=seeds[lo:hi], adapter_id=ad[lo:hi], request_base=lo, want_hashes=True, max_blocks=B)
        hashes = np.ascontiguousarray(res["hashes_out"][:, :B])
        g_pick = torch.empty(world * Rs, dtype=torch.int32)
        g_nh = torch.empty(world * Rs, dtype=torch.uint16)
        g_hash = torch.empty((world * Rs, B), dtype=torch.int64)
        sharding.gather_commit_stream(dist, torch.from_numpy(res["pick"]), torch.from_numpy(res["total_blocks"]),
                                      torch.from_numpy(hashes.view(np.int64)), g_pick, g_nh, g_hash)
        idx.commit(g_pick.numpy(), g_hash.numpy().view(np.uint64), g_nh.numpy())
        picks_seen.append(g_pick.numpy().copy())
        if rank == 0:                                     # the single-process run of the same unsharded batch
            whole = o.schedule_batch(osnap, prof, ref, world * Rs, prompt_bytes=prompts, prompt_off=off, model_seed=seeds,
                                     adapter_id=ad, want_hashes=True, max_blocks=B)
            assert np.array_equal(whole["pick"], g_pick.numpy()), f"batch {k}: sharded picks differ from the unsharded batch"
            ref.commit(whole["pick"], np.ascontiguousarray(whole["hashes_out"][:, :B]), whole["total_blocks"])
    # replicas identical: every rank's LRU contents, endpoint by endpoint, gathered on rank 0 and compared with the reference
    keys = [list(idx.lru_keys(m) or []) for m in range(Mc)]   # None: the endpoint never got an LRU
    flat = np.concatenate([np.array([len(x)], np.uint64) for x in keys] + [np.asarray(x, np.uint64) for x in keys])
    width = torch.tensor([len(flat)], dtype=torch.int64)
    widths = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(widths, width)
    same_len = all(int(w.item()) == len(flat) for w in widths)
    mine = torch.from_numpy(flat.view(np.int64).copy())
    allk = [torch.empty_like(mine) for _ in range(world)] if same_len else None
    if same_len:
        dist.all_gather(allk, mine)
    if rank == 0:
        identical = same_len and all(torch.equal(allk[0], allk[r]) for r in range(world))
        ref_keys = [list(ref.lru_keys(m) or []) for m in range(Mc)]
        matches_ref = all(list(keys[m]) == list(ref_keys[m]) for m in range(Mc))
        touched = sum(1 for m in range(Mc) if len(keys[m]) > 0)
        evicting = max(len(x) for x in keys) == 40
        q.put((bool(identical), bool(matches_ref), touched, bool(evicting)))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_batch():
    port = 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    same, max_ties = q.get(timeout=5)
    assert same and max_ties > 1


def test_two_rank_latency_weighted_random_is_shard_invariant():
    port = 31500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_latency_weighted_random, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    same, distinct = q.get(timeout=5)
    assert same and distinct > 20   # a stochastic picker spreads the picks, identically with and without sharding


def test_snapshot_pack_roundtrip_and_alignment():
    import importlib.util
    spec = importlib.util.spec_from_file_location("sharding", os.path.join(_pkg.PKG_DIR, "sharding.py"))
    sharding = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sharding)
    snap = synth_snapshot(77, A=130, seed=1)  # 3 LoRA words per endpoint
    buf, layout = sharding.pack_snapshot(snap)
    back = sharding.unpack_snapshot(buf, layout)
    for name, off, nbytes, dt, shape in layout:
        assert off % 16 == 0
        assert np.array_equal(back[name], np.asarray(snap[name]).reshape(shape))
    assert sharding.shard_range(10, 0, 3) == (0, 3) and sharding.shard_range(10, 2, 3) == (6, 10)


def test_two_rank_closed_loop_replicas_replay_the_gathered_commit_stream():
    port = 33500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_closed_loop, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    identical, matches_ref, touched, evicting = q.get(timeout=5)
    assert identical and matches_ref and touched >= 2 and evicting   # (some LRU is full: the replay also evicted)
