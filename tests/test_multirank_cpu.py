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
