"""BASELINE configs[4] sweep + configs[3] point: audio-s/s of the bf16 forward over batch x length
(CUDA events, CUDA-graph replay, 1 GPU).  Writes a markdown table to stdout."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cse_b200  # noqa: E402,F401
from cse_b200 import shapes, synth  # noqa: E402
from cse_b200.models.ContExt import Sepformer as ContExt  # noqa: E402
from cse_b200.models.ContSep import Sepformer as ContSep  # noqa: E402


def time_model(m, args, iters):
    with torch.no_grad():
        for _ in range(3):
            m(*args)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            m(*args)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = "cuda"
    m = ContSep(2, add_mt=True)
    m.add_mt_pipeline()
    m.load_state_dict(synth.make_state_dict("contsep", 2, seed=0))
    m = m.to(dev).eval()
    m.precision, m.use_cuda_graph = "bf16", True
    secs, batches = [2, 4, 8, 16, 32], [1, 2, 4, 8, 16, 32, 64]
    print("| B \\ seconds | " + " | ".join(f"{s} s" for s in secs) + " |")
    print("|---|" + "---|" * len(secs))
    for B in batches:
        row = []
        for s in secs:
            T = s * 8000
            ps = shapes.path_shape(B, T, 1, 2)
            mix = torch.randn(B, T, device=dev) * 0.1
            ctx = torch.randn(B, 1, 4096, device=dev)
            iters = max(2, min(20, int(2e13 / shapes.algorithmic_flops(ps))))
            ms = time_model(m, (mix, ctx), iters)
            tf = shapes.algorithmic_flops(ps) / ms / 1e9
            row.append(f"{B * s / ms * 1e3:.0f} ({ms:.1f} ms, {tf:.0f} TF/s)")
            m._graphs.clear()
            torch.cuda.empty_cache()
        print(f"| {B} | " + " | ".join(row) + " |", flush=True)
    # configs[3]: HContExt TEDLIUM3-shape 3-spk, 16 s, c = 2
    h = ContExt(3, add_ctx=True, add_se=True)
    h.add_ctx_pipeline()
    h.add_se_pipeline()
    h.load_state_dict(synth.make_state_dict("hcontext", 3, seed=0))
    h = h.to(dev).eval()
    h.precision, h.use_cuda_graph = "bf16", True
    print("\n| HContExt 3-spk 16 s (c=2, n_inter=132) | ms / forward | audio-s/s |\n|---|---|---|")
    for B in (1, 8):
        mix = torch.randn(B, 128000, device=dev) * 0.1
        ctx = torch.randn(B, 1, 4096, device=dev)
        se = torch.randn(B, 1, 192, device=dev)
        ms = time_model(h, (mix, ctx, se), 5)
        print(f"| B={B} | {ms:.2f} | {B * 16 / ms * 1e3:.0f} |", flush=True)


if __name__ == "__main__":
    main()
