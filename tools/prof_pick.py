"""Times pick_sparse alone at the headline shape for several diagnostics settings. usage: prof_pick.py [occ ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import _pkg  # noqa: E402
from tests.helpers import synth_prompts, synth_snapshot, zipf_adapters  # noqa: E402

pkg = _pkg.load()
R, M, A = 65536, bench.M, bench.A
snap = synth_snapshot(M, A=A, seed=1)
o, osnap, prof, idx, seed, warm = bench.oracle_setup(snap)
eng = pkg.Engine(pkg.default_config(bench.SCORERS, max_endpoints=M, max_adapters=A, block_chars=bench.BLOCK_CHARS,
                                    max_blocks=bench.MAX_BLOCKS, prefix_capacity=1 << 19))
eng.set_snapshot(**snap)
eng.commit_picks(warm["pick"], warm["hashes_out"], warm["total_blocks"])
dev = torch.device("cuda:0")
sets = []
for i in range(3):
    p, off, _ = synth_prompts(R, prompt_len=bench.PROMPT_LEN, groups=150, shared=1024, seed=100 + i, prefix_seed=7)
    h, n = eng.hash_prompts(p, off, np.full(R, seed, np.uint64), block_chars=bench.BLOCK_CHARS, max_blocks=bench.MAX_BLOCKS)
    sets.append((torch.from_numpy(h.view(np.int64)).to(dev), torch.from_numpy(n.view(np.int16)).to(dev),
                 torch.from_numpy(zipf_adapters(R, A=A, seed=100 + i)).to(dev)))
out = dict(pick=torch.empty(R, dtype=torch.int32, device=dev), pick_score=torch.empty(R, dtype=torch.float64, device=dev),
           tie_count=torch.empty(R, dtype=torch.int32, device=dev))
stream = torch.cuda.Stream(device=dev)
for occ in [int(x) for x in sys.argv[1:]] or [0]:
    if occ:
        eng.set_debug(3, occ)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 60
    for i in range(iters + 6):
        if i == 6:
            torch.cuda.synchronize()
            a.record(stream)
        hh, nn, ad = sets[i % 3]
        eng.schedule(R, hashes_in=hh, n_hashes_in=nn, hash_stride=bench.MAX_BLOCKS, adapter_id=ad, device=True,
                     stream=stream.cuda_stream, out=out)
    b.record(stream)
    torch.cuda.synchronize()
    print(f"setting {occ}: pick_sparse {a.elapsed_time(b) / iters * 1e3:.2f} us  picks_sum {int(out['pick'].sum())}")
