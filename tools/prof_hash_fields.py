"""Profiling driver (ncu target): hash_bodies / hash_chain / count_fields at the headline shape (64K x 2 KB prompts)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg  # noqa: E402
from tests.helpers import synth_prompts  # noqa: E402

pkg = _pkg.load()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
eng = pkg.Engine(pkg.default_config([("prefix", 1.0)], max_endpoints=1024))
prompts, off, _ = synth_prompts(R, prompt_len=2048, groups=150, shared=1024, seed=1)
seeds = np.full(R, eng.model_seed("m"), np.uint64)
for _ in range(2):
    h, n = eng.hash_prompts(prompts, off, seeds)
    c = eng.count_fields(prompts, off)
print("hashed", int(n.sum()), "blocks; fields", int(c.sum()))
