"""One launch of the fused out-proj + residual + LayerNorm kernel at the cfg 2 intra shape (for ncu)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cse_b200  # noqa: E402,F401
from cse_b200 import _lib  # noqa: E402

M, K = 136544, 256
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
A = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
W = (torch.randn(256, K, device=dev, generator=g) / 16).to(torch.bfloat16)
bias, gam, bet = torch.randn(256, device=dev), torch.ones(256, device=dev), torch.zeros(256, device=dev)
R = torch.randn(M, 256, device=dev, generator=g)
H = torch.empty(M, 256, dtype=torch.bfloat16, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    _lib.call("cse_linear_residual_ln", _lib.ptr(A), K, _lib.ptr(W), _lib.ptr(bias), _lib.ptr(R), _lib.ptr(gam),
              _lib.ptr(bet), 1e-6, _lib.ptr(H), M, K, st)
torch.cuda.synchronize()
print("ok")
