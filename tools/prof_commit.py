"""Micro-benchmark of the PreRequest kernel (index_commit_kernel): synthetic commit batches with a controlled distribution of
picks over the endpoints.  usage: python tools/prof_commit.py [R] [n_hot_endpoints ...]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg  # noqa: E402

_pkg.load_build().build()
pkg = _pkg.load()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
hots = [int(x) for x in sys.argv[2:]] or [1024, 64, 8]
M, NB = 1024, 32
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
rng = np.random.Generator(np.random.PCG64(1))
groups = rng.integers(1, 2 ** 63, size=(150, 16), dtype=np.uint64)
for hot in hots:
    eng = pkg.Engine(pkg.default_config([("prefix", 1.0)], max_endpoints=M, prefix_capacity=1 << 22))
    times = []
    for it in range(6):
        g = rng.integers(0, 150, size=R)
        hashes = np.empty((R, NB), np.uint64)
        hashes[:, :16] = groups[g]
        hashes[:, 16:] = rng.integers(1, 2 ** 63, size=(R, 16), dtype=np.uint64)
        pick = rng.integers(0, hot, size=R).astype(np.int32)
        dp, dh = torch.from_numpy(pick).to(dev), torch.from_numpy(hashes.view(np.int64)).to(dev)
        dn = torch.full((R,), NB, dtype=torch.int16, device=dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.commit_picks_device(dp, dh, dn, touch_bound=R * NB, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    st = eng.stats()
    print(f"picks over {hot:5d} endpoints: commit of {R} x {NB} per batch: " + " ".join(f"{t * 1e3:8.2f}" for t in times) +
          f" ms | {R / np.median(times[2:]) / 1e6:7.2f} M req/s | live {st.prefix_live_hashes} used {st.prefix_hashes} lru {st.lru_entries} "
          f"rebuilds {st.prefix_rebuilds} err {st.index_error}")
    eng.close()
