"""Stock-PyTorch (eager) GPU yardstick for the Sepformer hot path — NOT a product path.

SURVEY.md §2b / §7 step 0: the reference owns no kernel, so "the reference on a B200" is whatever stock
PyTorch dispatches for its modules — cuDNN/cuBLAS GEMMs, the bundled flash-SDPA under autocast, ATen
norm / elementwise kernels, eager launches.  This file runs that: the same op sequence as
`Sepformer.forward` (ContSep.py:53-100, ContExt.py:54-129), `Dual_Path_Model_CSE.forward`
(ContSep.py:205-268), `Dual_Computation_Block_CSE.forward` (ContSep.py:453-533) and
`SBTransformerBlock_CSE` / `TransformerEncoderLayer` / `MultiheadAttention` (CSE_transformer.py:90-106,
385-416, 535-557), composed from the stock `torch.nn` modules our module mirror already owns as parameter
containers (`nn.Conv1d`, `nn.GroupNorm`, `nn.MultiheadAttention`, `nn.LayerNorm`, `nn.Linear`, `nn.PReLU`,
`nn.Conv2d`, `nn.ConvTranspose1d`) in `oracle/eager_reference.py`.  None of our kernels, C ABI or engine is on
that path.

    python tools/eager_yardstick.py [--batch 16] [--seconds 4] [--steps 10] [--dtype bf16|fp32]
        -> one JSON line: audio-s/s of the eager path, ours beside it (same weights, same inputs), ours / eager.

tests/test_baseline_shapes_gpu.py uses the same module as the CUDA-autocast drift yardstick (the reference's real
`--bf16` arithmetic, train_ContSep.py:383), after checking its fp32 output against the oracle.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.eager_reference import run_eager  # noqa: E402  (the stock-torch restatement lives with the oracle)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--seconds", type=int, default=4)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16", "fp32"])
    args = ap.parse_args()
    import cse_b200  # noqa: F401
    from cse_b200 import synth
    from cse_b200.models.ContSep import Sepformer

    dev = torch.device("cuda", 0)
    T = args.seconds * 8000
    model = Sepformer(2, add_mt=True)
    model.add_mt_pipeline()
    model.load_state_dict(synth.make_state_dict("contsep", 2, seed=0))
    model = model.to(dev).eval()
    mix, _ = synth.make_mixture(args.batch, T, 2, seed=1234)
    ctx = synth.make_context(args.batch, 1, seed=1234)
    mix, ctx = mix.to(dev), ctx.to(dev)

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps, (time.perf_counter() - t0) / args.steps * 1e3

    eager_ms, eager_wall = timed(lambda: run_eager(model, mix, ctx, args.dtype))
    model.precision = "fp32" if args.dtype == "fp32" else "bf16"

    def ours():
        with torch.no_grad():
            return model(mix, ctx)

    ours_ms, _ = timed(ours)
    model.use_cuda_graph = True
    graph_ms, _ = timed(ours)
    est_e, _ = run_eager(model, mix, ctx, args.dtype)
    est_o, _ = ours()
    rel = ((est_o.double() - est_e.double()).norm() / est_e.double().norm()).item()
    audio = args.batch * args.seconds
    print(json.dumps({
        "yardstick": "stock PyTorch eager (cuBLAS/cuDNN + flash-SDPA under autocast + ATen), reference op sequence",
        "config": {"batch": args.batch, "seconds": args.seconds, "dtype": args.dtype, "model": "ContSep 2-spk c=1"},
        "eager_ms": eager_ms, "eager_wall_ms": eager_wall, "eager_audio_s_per_s": audio / (eager_ms / 1e3),
        "ours_eager_launch_ms": ours_ms, "ours_graph_ms": graph_ms, "ours_audio_s_per_s": audio / (graph_ms / 1e3),
        "ours_over_eager": eager_ms / graph_ms, "rel_l2_ours_vs_eager": rel,
        "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0)}), flush=True)


if __name__ == "__main__":
    main()
