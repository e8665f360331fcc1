"""Attention kernel A/B timing + pipeline trace of the tcgen05 v4 kernel (debug tool, not a product path).

    python tools/attn_trace.py            # on the GPU box

1. CUDA-event timing (20 launches after 3 warm-ups, inputs larger than L2 for the big shapes) of every bf16
   attention kernel at the shapes of BASELINE configs[1] / [3]:  mma.sync | tcgen05 v1 | tcgen05 v4 | v4 with 2/8 and 3/8 polynomial exponentials.
2. clock64 trace of CTA 0 of the v4 kernel (cse_debug_attention_trace) at the intra shape: per item, cycles relative
   to the item's S issue, so the dead time between the MUFU and tensor phases is visible.
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cse_b200  # noqa: E402,F401
from cse_b200 import _lib  # noqa: E402

DEV = "cuda:0"
TRACE_MODE = int(os.environ.get("ATTN_TRACE_MODE", "4"))
BF16 = _lib.BF16
SLOTS = ["S issue", "PV0 issue", "PV1", "PV2", "PV3", "sm wait S", "sm S seen", "sm 1st chunk", "sm 2nd chunk", "sm max/sum xchg",
         "-", "sm O seen", "sm epi done", "QK load issue", "V load issue", "-"]


def run(qkv, nseq, n, out, mode):
    lib = _lib.load()
    lib.cse_debug_force_mma_attention(mode)
    try:
        _lib.call("cse_attention_fwd", _lib.ptr(qkv), nseq, n, BF16, _lib.ptr(out),
                  C.c_void_p(torch.cuda.current_stream().cuda_stream))
    finally:
        lib.cse_debug_force_mma_attention(0)


def timeit(nseq, n, mode, iters=20):
    g = torch.Generator(device=DEV).manual_seed(1)
    qkv = (torch.randn(nseq * n, 768, device=DEV, generator=g) * 1.5).to(torch.bfloat16)
    out = torch.empty(nseq * n, 256, dtype=torch.bfloat16, device=DEV)
    for _ in range(3):
        run(qkv, nseq, n, out, mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run(qkv, nseq, n, out, mode)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3, out


def main():
    names = {1: "mma.sync", 2: "tcgen05 v1", 4: "tcgen05 v4"}
    modes = (1, 2, 4)
    print("shape (nseq x n)          " + "".join(f"{names[m]:>14}" for m in modes) + "   [us per launch]")
    for nseq, n in [(544, 251), (4000, 35), (1040, 252), (2000, 132), (4000, 67), (4000, 19)]:
        row, outs = [], {}
        for m in modes:
            try:
                us, outs[m] = timeit(nseq, n, m)
                row.append(f"{us:14.1f}")
            except Exception as e:  # noqa: BLE001
                row.append(f"{'ERR':>14}")
                print("   ", type(e).__name__, str(e)[:200])
        agree = ""
        for m in (4,):
            if 1 in outs and m in outs:
                d = (outs[m].float() - outs[1].float()).norm() / outs[1].float().norm()
                agree += f"   mode {m} vs mma rel-L2 {d.item():.2e}"
        print(f"{nseq:6d} x {n:3d}             " + "".join(row) + agree)

    # ---- trace of CTA 0, v4, intra shape ----
    nseq, n = 544, 251
    g = torch.Generator(device=DEV).manual_seed(1)
    qkv = (torch.randn(nseq * n, 768, device=DEV, generator=g) * 1.5).to(torch.bfloat16)
    out = torch.empty(nseq * n, 256, dtype=torch.bfloat16, device=DEV)
    buf = torch.zeros(64 * 16, dtype=torch.int64, device=DEV)
    lib = _lib.load()
    run(qkv, nseq, n, out, TRACE_MODE)
    torch.cuda.synchronize()
    lib.cse_debug_attention_trace(C.c_void_p(buf.data_ptr()))
    run(qkv, nseq, n, out, TRACE_MODE)
    torch.cuda.synchronize()
    lib.cse_debug_attention_trace(None)
    t = buf.cpu().view(64, 16)
    t0 = int(t[0][t[0] > 0].min())
    print(f"\nmode {TRACE_MODE} trace, CTA 0, intra shape 544 x 251: cycles since the CTA's first event; one line per item "
          "(even items = softmax group 0, odd = group 1)")
    print("item  " + "".join(f"{s:>14}" for s in SLOTS[:15]))
    for k in range(64):
        if int(t[k].max()) == 0:
            break
        print(f"{k:4d}  " + "".join(f"{(int(v) - t0) if int(v) > 0 else -1:14d}" for v in t[k][:15]))
    last = max(int(t[k][12]) for k in range(64) if int(t[k][12]) > 0)
    n_items = sum(1 for k in range(64) if int(t[k][12]) > 0)
    print(f"items traced {n_items}; cycles from first event to last epilogue {last - t0} = {(last - t0) / max(n_items, 1):.0f} per item")


if __name__ == "__main__":
    main()
