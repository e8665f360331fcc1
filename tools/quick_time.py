"""Scratch timing of the whole forward (CUDA events) — development aid, not the bench."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cse_b200  # noqa: E402,F401
from cse_b200 import synth  # noqa: E402
from cse_b200.models.ContSep import Sepformer  # noqa: E402


def main():
    cases = [(16, 32000, "bf16", 5), (1, 32000, "bf16", 5), (16, 32000, "fp32", 2), (4, 128000, "bf16", 3)]
    if len(sys.argv) > 1:
        cases = [(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]))]
    sd = synth.make_state_dict("contsep", 2, seed=0)
    m = Sepformer(2, add_mt=True)
    m.add_mt_pipeline()
    m.load_state_dict(sd)
    m = m.cuda().eval()
    m.use_cuda_graph = len(sys.argv) > 5 and sys.argv[5] == "graph"
    for B, T, prec, iters in cases:
        mix = torch.randn(B, T, device="cuda") * 0.1
        ctx = torch.randn(B, 1, 4096, device="cuda")
        m.precision = prec
        with torch.no_grad():
            for _ in range(2):
                m(mix, ctx)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(iters):
                est, _ = m(mix, ctx)
            e1.record()
            t_cpu = (time.perf_counter() - t0) / iters * 1e3
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"B={B} T={T} {prec}: {ms:.2f} ms/fwd  (cpu enqueue {t_cpu:.2f} ms)  "
              f"{B * T / 8000 / ms * 1e3:.0f} audio-s/s  finite={bool(torch.isfinite(est).all())}", flush=True)


if __name__ == "__main__":
    main()
