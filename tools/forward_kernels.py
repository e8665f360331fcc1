"""Kernel-time table of the cfg 2 forward (debug tool): torch.profiler over 3 eager bf16 forwards."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cse_b200  # noqa: E402,F401
from cse_b200 import synth  # noqa: E402
from cse_b200.models.ContSep import Sepformer  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

dev = "cuda:0"
m = Sepformer(2, add_mt=True)
m.add_mt_pipeline()
m.load_state_dict(synth.make_state_dict("contsep", 2, seed=0))
m = m.to(dev).eval()
m.use_cuda_graph = False
mix, _ = synth.make_mixture(16, 32000, 2, seed=1)
ctx = synth.make_context(16, 1, seed=1)
mix, ctx = mix.to(dev), ctx.to(dev)
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    for _ in range(3):
        m(mix, ctx)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            m(mix, ctx)
        torch.cuda.synchronize()
rows = [(e.key, e.count, getattr(e, "device_time_total", getattr(e, "cuda_time_total", 0))) for e in prof.key_averages()]
rows = [r for r in rows if r[2] > 0]
tot = sum(r[2] for r in rows)
print(f"GPU kernel time per forward: {tot / 3 / 1e3:.2f} ms")
for k, c, t in sorted(rows, key=lambda r: -r[2]):
    print(f"{k[:80]:82s} {c / 3:6.1f} /fwd {t / 3:9.1f} us/fwd {t / c:8.1f} us each {100 * t / tot:5.1f}%")
