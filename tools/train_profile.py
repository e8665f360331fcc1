"""Kernel-time table of one autocast training step (debug tool, not a product path): torch.profiler over 3 steps of
the `bench.py --workload train` step (ContExt 2-spk, 2 x 4 s, bf16 autocast, fused optimiser).

    python tools/train_profile.py > gpurun_out/train_kernels.txt
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cse_b200  # noqa: E402,F401
from cse_b200 import losses, synth  # noqa: E402
from cse_b200.models.ContExt import Sepformer  # noqa: E402
from cse_b200.optim import AdamW  # noqa: E402

dev = "cuda:0"
model = Sepformer(2, add_ctx=True)
model.add_ctx_pipeline()
model.load_state_dict(synth.make_state_dict("context", 2, seed=0))
model = model.to(dev).train()
opt = AdamW(model.parameters(), lr=1e-4, amsgrad=True)
sisnr = losses.ScaleInvariantSignalNoiseRatio()
mix, src = synth.make_mixture(2, 32000, 2, seed=4321)
ctx = synth.make_context(2, 1, seed=4321)
mix, ctx, tgt = mix.to(dev), ctx.to(dev), src[:, :, 0].contiguous().to(dev)


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        est = model(mix, ctx)
        loss = -sisnr(est[:, :, 0], tgt)
    loss.backward()
    opt.step(max_norm=5.0)
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print(f"wall per step (no profiler): {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms")
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
ka = prof.key_averages()
rows = [(e.key, e.count, getattr(e, "device_time_total", getattr(e, "cuda_time_total", 0))) for e in ka
        if getattr(e, "device_type", None) is not None and str(e.device_type).endswith("CUDA")]
if not rows:
    rows = [(e.key, e.count, getattr(e, "self_device_time_total", getattr(e, "self_cuda_time_total", 0))) for e in ka]
    rows = [r for r in rows if r[2] > 0]
tot = sum(r[2] for r in rows)
print(f"GPU kernel time per step: {tot / 3 / 1e3:.2f} ms over {sum(r[1] for r in rows) / 3:.0f} launches")
for k, c, t in sorted(rows, key=lambda r: -r[2])[:40]:
    print(f"{k[:90]:92s} {c / 3:8.1f} /step {t / 3:10.1f} us/step {t / c:9.1f} us each {100 * t / tot:5.1f}%")
