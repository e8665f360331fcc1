"""Per-step device time of the autocast training step (debug tool): is the step time stable over many un-synchronised steps?"""
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
T = int(sys.argv[1]) if len(sys.argv) > 1 else 32000
model = Sepformer(2, add_ctx=True)
model.add_ctx_pipeline()
model.load_state_dict(synth.make_state_dict("context", 2, seed=0))
model = model.to(dev).train()
opt = AdamW(model.parameters(), lr=1e-4, amsgrad=True)
sisnr = losses.ScaleInvariantSignalNoiseRatio()
mix, src = synth.make_mixture(2, T, 2, seed=4321)
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
n = 24
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
cpu = []
ev[0].record()
for i in range(n):
    t0 = time.perf_counter()
    step()
    cpu.append((time.perf_counter() - t0) * 1e3)
    ev[i + 1].record()
torch.cuda.synchronize()
gpu = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
print("T", T, "reserved MB", torch.cuda.memory_reserved() >> 20, "allocated MB", torch.cuda.memory_allocated() >> 20)
print("gpu ms per step:", " ".join(f"{x:.1f}" for x in gpu))
print("cpu ms per step:", " ".join(f"{x:.1f}" for x in cpu))
print("retries", torch.cuda.memory_stats().get("num_alloc_retries"), "cudaMalloc calls", torch.cuda.memory_stats().get("num_device_alloc"))
