"""GPU: the tcgen05/TMA bf16 GEMM against a float64 product of the same bf16-rounded operands."""
import ctypes as C
import math

import pytest
import torch

import cse_b200  # noqa: F401
from cse_b200 import _lib
from cse_b200._lib import BF16
from helpers import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _rand(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


SHAPES = [(128, 256, 256), (1, 256, 256), (130, 768, 256), (1000, 1024, 256), (517, 256, 1024),
          (4099, 512, 256), (300, 128, 64), (20000, 768, 256), (2 * 148 * 128 + 77, 256, 256)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_tc_epilogues(M, N, K):
    A = _rand(M, K, seed=1).to(torch.bfloat16)
    W = (_rand(N, K, seed=2) / math.sqrt(K)).to(torch.bfloat16)
    bias = _rand(N, seed=3)
    res = _rand(M, N, seed=4)
    Ad, Wd, bd = A.to(DEV), W.to(DEV), bias.to(DEV)
    exact = A.double() @ W.double().t()
    # bf16 out, bias
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    _lib.call("cse_linear", _lib.ptr(Ad), K, _lib.ptr(Wd), _lib.ptr(bd), 1.0, None, _lib.ptr(out), N,
              M, N, K, 0, 0, BF16, _st())
    torch.cuda.synchronize()
    assert rel_l2(out.float().cpu(), exact + bias.double()) < 4e-3          # bf16 output rounding
    # bf16 out, relu, doubled bias
    _lib.call("cse_linear", _lib.ptr(Ad), K, _lib.ptr(Wd), _lib.ptr(bd), 2.0, None, _lib.ptr(out), N,
              M, N, K, 1, 0, BF16, _st())
    assert rel_l2(out.float().cpu(), torch.relu(exact + 2 * bias.double())) < 4e-3
    # fp32 out, no bias
    o32 = torch.empty(M, N, dtype=torch.float32, device=DEV)
    _lib.call("cse_linear", _lib.ptr(Ad), K, _lib.ptr(Wd), None, 0.0, None, _lib.ptr(o32), N,
              M, N, K, 0, 1, BF16, _st())
    assert rel_l2(o32.cpu(), exact) < 1e-5                                   # fp32 accumulate
    # fp32 in-place residual
    r = res.to(DEV).clone()
    _lib.call("cse_linear", _lib.ptr(Ad), K, _lib.ptr(Wd), _lib.ptr(bd), 1.0, _lib.ptr(r), _lib.ptr(r), N,
              M, N, K, 0, 1, BF16, _st())
    assert rel_l2(r.cpu(), exact + bias.double() + res.double()) < 1e-5


@pytest.mark.parametrize("M,N,relu", [(1, 256, 0), (128, 768, 0), (129, 768, 0), (517, 1024, 1), (20000, 768, 0),
                                      (2 * 148 * 128 + 77, 1024, 1), (3 * 148 * 128 + 5, 256, 0)])
def test_ln_fused_gemm(M, N, relu):
    """cse_ln_linear: LayerNorm in the A-operand producer == LayerNorm kernel -> bf16 -> GEMM."""
    R = _rand(M, 256, seed=7) * 2.0 + 0.3
    g, b = 1 + 0.1 * _rand(256, seed=8), 0.1 * _rand(256, seed=9)
    W = (_rand(N, 256, seed=10) / 16).to(torch.bfloat16)
    bias = _rand(N, seed=11)
    Rd, gd, bd, Wd, biasd = R.to(DEV), g.to(DEV), b.to(DEV), W.to(DEV), bias.to(DEV)
    out = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
    _lib.call("cse_ln_linear", _lib.ptr(Rd), _lib.ptr(gd), _lib.ptr(bd), 1e-6, _lib.ptr(Wd), _lib.ptr(biasd),
              _lib.ptr(out), N, M, N, relu, _st())
    torch.cuda.synchronize()
    x = R.double()
    ln = (x - x.mean(-1, keepdim=True)) / torch.sqrt(x.var(-1, unbiased=False, keepdim=True) + 1e-6) * g.double() + b.double()
    ref = ln.float().to(torch.bfloat16).double() @ W.double().t() + bias.double()   # same bf16 rounding of A
    if relu:
        ref = torch.relu(ref)
    assert rel_l2(out.float().cpu(), ref) < 5e-3


@pytest.mark.parametrize("M,K", [(1, 256), (128, 256), (129, 256), (517, 256), (20000, 256), (2 * 148 * 128 + 77, 256),
                                 (136544, 256), (1000, 64), (777, 1024)])
def test_linear_residual_ln(M, K):
    """cse_linear_residual_ln == out-proj GEMM + residual add + LayerNorm kernel (`src = src + att; norm2(src)`,
    CSE_transformer.py:399-408): R against fp64, H against fp64 LayerNorm of OUR fp32 R (bf16 output rounding)."""
    A = _rand(M, K, seed=31).to(torch.bfloat16)
    W = (_rand(256, K, seed=32) / (K ** 0.5)).to(torch.bfloat16)
    bias, R0 = _rand(256, seed=33), _rand(M, 256, seed=34) * 2.0 + 0.5
    g, b = 1 + 0.1 * _rand(256, seed=35), 0.1 * _rand(256, seed=36)
    Ad, Wd, biasd, gd, bd = A.to(DEV), W.to(DEV), bias.to(DEV), g.to(DEV), b.to(DEV)
    Rd = R0.to(DEV).clone()
    H = torch.zeros(M, 256, dtype=torch.bfloat16, device=DEV)
    _lib.call("cse_linear_residual_ln", _lib.ptr(Ad), K, _lib.ptr(Wd), _lib.ptr(biasd), _lib.ptr(Rd), _lib.ptr(gd),
              _lib.ptr(bd), 1e-6, _lib.ptr(H), M, K, _st())
    torch.cuda.synchronize()
    ref_R = R0.double() + A.double() @ W.double().t() + bias.double()
    assert rel_l2(Rd.cpu(), ref_R) < 2e-6
    x = Rd.cpu().double()
    ln = (x - x.mean(-1, keepdim=True)) / torch.sqrt(x.var(-1, unbiased=False, keepdim=True) + 1e-6) * g.double() + b.double()
    assert rel_l2(H.float().cpu(), ln) < 3e-3                                  # bf16 rounding of the output only
    assert (H.float().cpu() - ln).abs().max() < 0.04


@pytest.mark.parametrize("M", [1, 128, 129, 517, 20000, 2 * 148 * 128 + 77, 3 * 148 * 128 + 5])
def test_ffn_fused(M):
    """cse_ffn_fused == Linear(256,1024) -> ReLU -> bf16 -> Linear(1024,256) -> residual add
    (PositionalwiseFeedForward + `src = src + ...`, CSE_transformer.py:407-411,547-566)."""
    A = (_rand(M, 256, seed=21)).to(torch.bfloat16)
    W1 = (_rand(1024, 256, seed=22) / 16).to(torch.bfloat16)
    W2 = (_rand(256, 1024, seed=23) / 32).to(torch.bfloat16)
    b1, b2 = _rand(1024, seed=24), _rand(256, seed=25)
    R0 = _rand(M, 256, seed=26)
    Ad, W1d, W2d, b1d, b2d = A.to(DEV), W1.to(DEV), W2.to(DEV), b1.to(DEV), b2.to(DEV)
    Rd = R0.to(DEV).clone()
    _lib.call("cse_ffn_fused", _lib.ptr(Ad), _lib.ptr(W1d), _lib.ptr(b1d), _lib.ptr(W2d), _lib.ptr(b2d),
              _lib.ptr(Rd), M, _st())
    torch.cuda.synchronize()
    hid = torch.relu(A.double() @ W1.double().t() + b1.double()).float().to(torch.bfloat16).double()  # same rounding
    ref = R0.double() + hid @ W2.double().t() + b2.double()
    # 5e-5: a hidden value that lands within fp32-accumulation error of a bf16 rounding boundary may round
    # the other way than the float64 reference (one bf16 ulp on that element)
    assert rel_l2(Rd.cpu(), ref) < 5e-5
    # and against the two-GEMM path of the same library (bit-level differences only from accumulation order)
    F1 = torch.empty(M, 1024, dtype=torch.bfloat16, device=DEV)
    R2 = R0.to(DEV).clone()
    _lib.call("cse_linear", _lib.ptr(Ad), 256, _lib.ptr(W1d), _lib.ptr(b1d), 1.0, None, _lib.ptr(F1), 1024, M, 1024,
              256, 1, 0, BF16, _st())
    _lib.call("cse_linear", _lib.ptr(F1), 1024, _lib.ptr(W2d), _lib.ptr(b2d), 1.0, _lib.ptr(R2), _lib.ptr(R2), 256, M,
              256, 1024, 0, 1, BF16, _st())
    torch.cuda.synchronize()
    assert rel_l2(Rd.cpu(), R2.double().cpu()) < 5e-5


@pytest.mark.parametrize("M", [1, 128, 129, 517, 20000, 2 * 148 * 128 + 77, 3 * 148 * 128 + 5])
@pytest.mark.parametrize("with_next", [True, False])
def test_ffn_ln_fused(M, with_next):
    """cse_ffn_ln_fused == layernorm_kernel (norm2) -> cse_ffn_fused -> layernorm_kernel (next layer's norm1)
    (CSE_transformer.py:406-411 and :387 of the following layer).  The LayerNorm warps run layernorm_kernel's
    arithmetic and the GEMMs see the same operands, so the three results are BIT-EXACT with the three-launch path;
    the float64 check guards that path itself."""
    W1 = (_rand(1024, 256, seed=22) / 16).to(torch.bfloat16).to(DEV)
    W2 = (_rand(256, 1024, seed=23) / 32).to(torch.bfloat16).to(DEV)
    b1, b2 = _rand(1024, seed=24).to(DEV), _rand(256, seed=25).to(DEV)
    g2, be2 = (1 + 0.1 * _rand(256, seed=27)).to(DEV), (0.1 * _rand(256, seed=28)).to(DEV)
    g1, be1 = (1 + 0.1 * _rand(256, seed=29)).to(DEV), (0.1 * _rand(256, seed=30)).to(DEV)
    R0 = (_rand(M, 256, seed=26) * 1.5 + 0.25).to(DEV)

    # three-launch path
    A_ref = torch.empty(M, 256, dtype=torch.bfloat16, device=DEV)
    R_ref = R0.clone()
    H_ref = torch.empty(M, 256, dtype=torch.bfloat16, device=DEV)
    _lib.call("cse_layernorm_fwd", _lib.ptr(R_ref), _lib.ptr(g2), _lib.ptr(be2), M, 1e-6, BF16, _lib.ptr(A_ref), _st())
    _lib.call("cse_ffn_fused", _lib.ptr(A_ref), _lib.ptr(W1), _lib.ptr(b1), _lib.ptr(W2), _lib.ptr(b2),
              _lib.ptr(R_ref), M, _st())
    _lib.call("cse_layernorm_fwd", _lib.ptr(R_ref), _lib.ptr(g1), _lib.ptr(be1), M, 1e-6, BF16, _lib.ptr(H_ref), _st())

    R = R0.clone()
    A = torch.full((M, 256), float("nan"), dtype=torch.bfloat16, device=DEV)
    H = torch.full((M, 256), float("nan"), dtype=torch.bfloat16, device=DEV)
    _lib.call("cse_ffn_ln_fused", _lib.ptr(R), _lib.ptr(g2), _lib.ptr(be2), 1e-6, _lib.ptr(A), _lib.ptr(W1),
              _lib.ptr(b1), _lib.ptr(W2), _lib.ptr(b2), _lib.ptr(g1) if with_next else None,
              _lib.ptr(be1) if with_next else None, _lib.ptr(H) if with_next else None, M, _st())
    torch.cuda.synchronize()
    assert torch.equal(A, A_ref)
    assert torch.equal(R, R_ref)
    if with_next:
        assert torch.equal(H, H_ref)
    else:
        assert torch.isnan(H.float()).all()

    x = R0.double().cpu()
    ln = lambda t, g, b: (t - t.mean(-1, keepdim=True)) / torch.sqrt(t.var(-1, unbiased=False, keepdim=True) + 1e-6) \
        * g.double().cpu() + b.double().cpu()  # noqa: E731
    a = ln(x, g2, be2).float().to(torch.bfloat16).double()
    hid = torch.relu(a @ W1.double().cpu().t() + b1.double().cpu()).float().to(torch.bfloat16).double()
    ref = x + hid @ W2.double().cpu().t() + b2.double().cpu()
    assert rel_l2(R.cpu(), ref) < 3e-4     # a LayerNorm output within rounding error of a bf16 boundary flips one ulp
    if with_next:
        assert rel_l2(H.float().cpu(), ln(R.double().cpu(), g1, be1)) < 3e-3


def test_gemm_tc_strided_views():
    """conv2d output [B*L, spk*256] re-read as [B*L*spk, 256] (abi.cu masknet_impl) and lda > K."""
    M, K = 777, 256
    big = _rand(M, 3 * K, seed=5).to(torch.bfloat16).to(DEV)
    W = (_rand(256, K, seed=6) / 16).to(torch.bfloat16).to(DEV)
    out = torch.empty(M, 256, dtype=torch.float32, device=DEV)
    A = big[:, K:2 * K]
    _lib.call("cse_linear", C.c_void_p(A.data_ptr()), 3 * K, _lib.ptr(W), None, 0.0, None, _lib.ptr(out), 256,
              M, 256, K, 0, 1, BF16, _st())
    assert rel_l2(out.cpu(), A.double().cpu() @ W.double().cpu().t()) < 1e-5


def test_gemm_tc_rejects_bad_shapes():
    a = torch.zeros(8, 100, dtype=torch.bfloat16, device=DEV)
    w = torch.zeros(256, 100, dtype=torch.bfloat16, device=DEV)
    o = torch.zeros(8, 256, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(_lib.CseError):
        _lib.call("cse_linear", _lib.ptr(a), 100, _lib.ptr(w), None, 0.0, None, _lib.ptr(o), 256, 8, 256, 100,
                  0, 0, BF16, _st())
