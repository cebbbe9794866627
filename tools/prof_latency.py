"""Profiling driver (ncu target): the latency-scorer profile at the headline shape, device-resident inputs.
usage: python tools/prof_latency.py [R] [M] [iters]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg  # noqa: E402
from tests.helpers import synth_snapshot  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
masked = len(sys.argv) > 4 and "masked" in sys.argv[4]
streaming = 0 if (len(sys.argv) > 4 and "nostream" in sys.argv[4]) else 1
pkg = _pkg.load()
coef = dict(ttft_intercept=12.5, ttft_kv=80.0, ttft_input=0.031, ttft_waiting=7.25, ttft_running=1.5, ttft_prefix=-40.0,
            tpot_intercept=9.0, tpot_kv=11.0, tpot_input=0.0007, tpot_waiting=0.9, tpot_running=0.35, tpot_generated=0.01,
            streaming_mode=streaming)
eng = pkg.Engine(pkg.default_config([("latency", 1.0)], max_endpoints=M, tie_mode=1, tie_seed=11))
eng.set_latency_params(pkg.latency_params(**coef))
rng = np.random.Generator(np.random.PCG64(77))
sd = synth_snapshot(M, seed=1)
eng.set_snapshot(sd["kv_usage"], sd["queue"], sd["running"], min_tpot_slo=rng.choice([0.0, 0.0, 22.0, 26.5, 60.0], M),
                 dispatched=rng.integers(0, 3, M).astype(np.int32), prefill_role=(rng.random(M) < 0.1).astype(np.uint8))
dev = torch.device("cuda:0")
req = dict(input_tokens=torch.from_numpy(rng.integers(16, 6000, R).astype(np.int32)).to(dev),
           ttft_slo=torch.from_numpy(rng.choice([0.0, 90.0, 140.0, 200.0, 400.0, 1e6], R)).to(dev),
           tpot_slo=torch.from_numpy(rng.choice([0.0, 18.0, 24.0, 30.0, 80.0], R)).to(dev))
if masked:
    req["cand_mask"] = torch.from_numpy(rng.integers(0, 2 ** 32, (R, (M + 31) // 32), dtype=np.uint64).astype(np.uint32).view(np.int32)).to(dev)
out = dict(pick=torch.empty(R, dtype=torch.int32, device=dev), pick_score=torch.empty(R, dtype=torch.float64, device=dev),
           tie_count=torch.empty(R, dtype=torch.int32, device=dev))
stream = torch.cuda.Stream(device=dev)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(iters + 1):
    if i == 1:
        torch.cuda.synchronize()
        a.record(stream)
    eng.schedule(R, device=True, stream=stream.cuda_stream, out=out, **req)
b.record(stream)
torch.cuda.synchronize()
print(f"R={R} M={M} masked={masked} streaming={streaming}: {a.elapsed_time(b) / iters * 1e3:.1f} us per launch, "
      f"{R / (a.elapsed_time(b) / iters * 1e-3) / 1e6:.1f} M picks/s; picked {(out['pick'] >= 0).sum().item()}")
