"""Development aid: the feed-forward kernel with and without its LayerNorm warps, in isolation (CUDA events, cfg 2 rows).
flags of cse_debug_ffn_ln: 1 = the norm2 warps skip their rows (results wrong, timing only), 2 = no next-layer norm1."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cse_b200  # noqa: E402,F401
from cse_b200 import _lib  # noqa: E402
from cse_b200._lib import BF16  # noqa: E402

DEV = "cuda:0"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 136544
lib = _lib.load()
lib.cse_debug_ffn_ln.argtypes = [C.c_int]
lib.cse_debug_ffn_ln.restype = None
g = torch.Generator().manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
W1 = (r(1024, 256) / 16).to(torch.bfloat16).to(DEV)
W2 = (r(256, 1024) / 32).to(torch.bfloat16).to(DEV)
b1, b2 = r(1024).to(DEV), r(256).to(DEV)
g2, be2, g1, be1 = (1 + 0.1 * r(256)).to(DEV), (0.1 * r(256)).to(DEV), (1 + 0.1 * r(256)).to(DEV), (0.1 * r(256)).to(DEV)
R = r(M, 256).to(DEV)
A = torch.zeros(M, 256, dtype=torch.bfloat16, device=DEV)
H = torch.zeros(M, 256, dtype=torch.bfloat16, device=DEV)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
p = _lib.ptr


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        R.mul_(0.5)                       # keeps R bounded; leaves its tail in L2 like out-proj does
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


def ln(x, gg, bb, out):
    _lib.call("cse_layernorm_fwd", p(x), p(gg), p(bb), M, 1e-6, BF16, p(out), st())


def ffn():
    _lib.call("cse_ffn_fused", p(A), p(W1), p(b1), p(W2), p(b2), p(R), M, st())


def ffn_ln(with_next):
    _lib.call("cse_ffn_ln_fused", p(R), p(g2), p(be2), 1e-6, p(A), p(W1), p(b1), p(W2), p(b2),
              p(g1) if with_next else None, p(be1) if with_next else None, p(H) if with_next else None, M, st())


print(f"M = {M}")
print(f"layernorm_kernel                     {timed(lambda: ln(R, g2, be2, A)):8.1f} us")
print(f"ffn_tc_kernel<false>                 {timed(ffn):8.1f} us")
print(f"LN2 + FFN + LN1 (three launches)     {timed(lambda: (ln(R, g2, be2, A), ffn(), ln(R, g1, be1, H))):8.1f} us")
for flags, what in [(0, "norm2 + norm1'"), (2, "norm2 only (flag)"), (1, "norm1' only (norm2 warps idle)"), (3, "neither (framework only)"),
                    (8, "both, E1 without bias loads"), (16, "both, wait_group.read (no IVALL)"), (24, "both, no bias, no IVALL")]:
    lib.cse_debug_ffn_ln(flags)
    print(f"ffn_tc_kernel<true> {what:32s} {timed(lambda: ffn_ln(True)):8.1f} us")
lib.cse_debug_ffn_ln(0)
print(f"ffn_tc_kernel<true> H1 = NULL                        {timed(lambda: ffn_ln(False)):8.1f} us")

# per-tile timeline of block 0 (flag 4)
lib.cse_debug_ffn_ln(4)
ffn_ln(True)
torch.cuda.synchronize()
lib.cse_debug_ffn_ln(0)
buf = (C.c_ulonglong * 192)()
lib.cse_debug_ffn_ln_trace.argtypes = [C.c_void_p]
lib.cse_debug_ffn_ln_trace(buf)
t = list(buf)
t0 = min(x for x in t if x)
rel = lambda i: (t[i] - t0) if t[i] else -1  # noqa: E731
print("tile:  Y ready | drained | reduce-adds complete | norm1 done || norm2 delivered | producer got rows   (cycles from the first stamp)")
for it in range(9):
    print(f"{it:4d}: {rel(4 * it):8d} {rel(4 * it + 1):8d} {rel(4 * it + 2):8d} {rel(4 * it + 3):8d} || {rel(128 + it):8d} {rel(160 + it):8d}")
