"""GPU: the performance-mode (bf16 tensor-core) backward — nn.Linear dgrad / wgrad on the tcgen05 GEMM
(csrc/backward_tc.cu, split-K wgrad), the mma.sync attention backward (csrc/attention_bwd_mma.cu) and one whole
transformer layer (cse_layer_bwd_bf16), i.e. what `loss.backward()` runs under torch.autocast
(train_ContSep.py:383-400).  Tolerance: bf16 operand rounding (2^-9 relative per element) against the fp64 closed
forms, i.e. ~4e-3 relative L2 for these reduction lengths.
"""
import ctypes as C

import pytest
import torch

import cse_b200  # noqa: F401
from cse_b200 import _lib
from helpers import rel_l2
from oracle import backward_oracle as BO

pytestmark = [pytest.mark.gpu]
DEV = "cuda:0"


def _rand(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize("M,N,K", [(300, 768, 256), (1000, 256, 1024), (5000, 1024, 256), (17068, 256, 256)])
@pytest.mark.parametrize("a_bf16", [False, True])
def test_linear_backward_tensor_core(M, N, K, a_bf16):
    a, w, dc = _rand(M, K, seed=1), _rand(N, K, seed=2) / K ** 0.5, _rand(M, N, seed=3)
    a_in = a.to(torch.bfloat16) if a_bf16 else a
    da_ref, dw_ref, db_ref = BO.manual_linear_bwd(a_in.double(), w.double(), dc.double())
    ad, wd, dcd = a_in.to(DEV), w.to(DEV), dc.to(DEV)
    da = torch.empty(M, K, device=DEV)
    dw, db = torch.zeros(N, K, device=DEV), torch.zeros(N, device=DEV)
    nbytes = _lib.load().cse_linear_bwd_tc_scratch_bytes(M, N, K)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.call("cse_linear_bwd_tc", _lib.ptr(ad), int(a_bf16), K, _lib.ptr(wd), _lib.ptr(dcd), M, N, K, _lib.ptr(da),
              1, K, _lib.ptr(dw), _lib.ptr(db), _lib.ptr(ws), nbytes, st)
    assert rel_l2(da.cpu(), da_ref) < 6e-3
    assert rel_l2(dw.cpu(), dw_ref) < 6e-3
    assert rel_l2(db.cpu(), db_ref) < 1e-4
    _lib.call("cse_linear_bwd_tc", _lib.ptr(ad), int(a_bf16), K, _lib.ptr(wd), _lib.ptr(dcd), M, N, K, None,
              1, K, _lib.ptr(dw), None, _lib.ptr(ws), nbytes, st)            # dW accumulates
    assert rel_l2(dw.cpu(), 2 * dw_ref) < 6e-3


@pytest.mark.parametrize("nseq,n", [(3, 35), (2, 251), (5, 1), (7, 16), (4, 64), (3, 65), (2, 132), (1, 256), (68, 251),
                                    (500, 35)])
def test_attention_backward_tensor_core(nseq, n):
    """cse_attention_bwd_bf16 (mma.sync, both orientations) against the fp64 closed form evaluated on the SAME bf16
    q/k/v; its `out` input is our own bf16 forward.  Per-matrix relative L2 <= 1.5e-2 (P, dS and dO are rounded to
    bf16 for the contractions)."""
    qkv = (_rand(nseq, n, 768, seed=9) * 1.2).to(torch.bfloat16)
    do = _rand(nseq, n, 256, seed=10)
    o_ref, dqkv_ref = BO.manual_attention_bwd(qkv.double(), do.double())
    qd, dod = qkv.to(DEV), do.to(DEV)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = torch.empty(nseq * n, 256, dtype=torch.bfloat16, device=DEV)
    _lib.call("cse_attention_fwd", _lib.ptr(qd), nseq, n, _lib.BF16, _lib.ptr(out), st)
    assert rel_l2(out.float().cpu().view(nseq, n, 256), o_ref) < 1e-2
    dqkv = torch.full((nseq, n, 768), float("nan"), device=DEV)
    _lib.call("cse_attention_bwd_bf16", _lib.ptr(qd), _lib.ptr(out), _lib.ptr(dod), nseq, n, _lib.ptr(dqkv), st)
    assert torch.isfinite(dqkv).all()
    for name, sl in (("dq", slice(0, 256)), ("dk", slice(256, 512)), ("dv", slice(512, 768))):
        got, ref = dqkv.cpu()[..., sl].double(), dqkv_ref[..., sl]
        # n == 1: softmax of a single score is constant, dq = dk = 0 — here up to the bf16 rounding of dP - D, which is
        # relative to dv: the floor is a fraction of the whole gradient's norm
        assert (got - ref).norm() <= 1.5e-2 * max(ref.norm().item(), 0.5 * dqkv_ref.norm().item()), name
    with pytest.raises(_lib.CseError):
        _lib.call("cse_attention_bwd_bf16", _lib.ptr(qd), _lib.ptr(out), _lib.ptr(dod), 1, 300, _lib.ptr(dqkv), st)


@pytest.mark.parametrize("nseq,n", [(4, 36), (3, 251)])
def test_layer_backward_performance_mode(nseq, n):
    """cse_layer_bwd_bf16 against autograd over the fp64 oracle, at the tolerance of bf16 operands: the
    CPU model of this composition (tests/test_backward_oracle.py::test_performance_mode_layer_backward_tolerance_
    is_reachable) puts one layer at 2-5 % per gradient; the bound is 8 %."""
    from cse_b200 import backward
    from test_backward_oracle import _layer_params
    p64 = _layer_params(71)
    x, dy = _rand(nseq, n, 256, seed=72), _rand(nseq, n, 256, seed=73)
    _, dx_ref, g_ref = BO.autograd_layer(p64, x.double(), dy.double())
    params = {k: v.float().to(DEV) for k, v in p64.items()}
    dR, grads = backward.layer_backward(params, x.reshape(nseq * n, 256).to(DEV),
                                        dy.reshape(nseq * n, 256).to(DEV), nseq, n, experimental_bf16=True)
    assert rel_l2(dR.cpu().view(nseq, n, 256), dx_ref) < 8e-2
    for k in BO.LAYER_KEYS:
        assert rel_l2(grads[k].cpu(), g_ref[k]) < 8e-2, k
