"""CPU: the closed-form gradients the CUDA backward kernels implement (oracle/backward_oracle.py
`manual_*`) against autograd over the pinned forward oracle — and, where /root/reference exists,
against autograd over the reference's own TransformerEncoderLayer."""
import os

import pytest
import torch

from helpers import rel_l2
from oracle import backward_oracle as BO
from oracle import sepformer_oracle as O
from cse_b200 import synth


def _layer_params(seed, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    shapes = {"self_att.att.in_proj_weight": (768, 256), "self_att.att.in_proj_bias": (768,),
              "self_att.att.out_proj.weight": (256, 256), "self_att.att.out_proj.bias": (256,),
              "pos_ffn.ffn.0.weight": (1024, 256), "pos_ffn.ffn.0.bias": (1024,),
              "pos_ffn.ffn.3.weight": (256, 1024), "pos_ffn.ffn.3.bias": (256,),
              "norm1.norm.weight": (256,), "norm1.norm.bias": (256,),
              "norm2.norm.weight": (256,), "norm2.norm.bias": (256,)}
    p = {}
    for k, s in shapes.items():
        if k.endswith("norm.weight"):
            p[k] = (1 + 0.1 * torch.randn(s, generator=g)).to(dtype)
        elif len(s) == 1:
            p[k] = (0.1 * torch.randn(s, generator=g)).to(dtype)
        else:
            p[k] = (torch.randn(s, generator=g) / s[1] ** 0.5).to(dtype)
    return p


@pytest.mark.parametrize("Bp,n", [(3, 7), (2, 37)])
def test_manual_layer_backward_matches_autograd(Bp, n):
    p = _layer_params(11)
    g = torch.Generator().manual_seed(12)
    x = torch.randn(Bp, n, 256, generator=g).double()
    dy = torch.randn(Bp, n, 256, generator=g).double()
    _, dx_ref, g_ref = BO.autograd_layer(p, x, dy)
    dx, grads = BO.manual_layer_bwd(p, x, dy)
    assert rel_l2(dx, dx_ref) < 1e-12
    for k in BO.LAYER_KEYS:
        assert rel_l2(grads[k], g_ref[k]) < 1e-12, k


def test_manual_attention_forward_is_the_oracle_attention():
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(2, 19, 768, generator=g).double()
    o, _ = BO.manual_attention_bwd(qkv, torch.zeros(2, 19, 256).double())
    sd = {"att.in_proj_weight": torch.eye(768, 256).double(), "att.in_proj_bias": torch.zeros(768).double(),
          "att.out_proj.weight": torch.eye(256).double(), "att.out_proj.bias": torch.zeros(256).double()}
    # oracle MHA on an input whose projection IS qkv: feed qkv through identity-free path instead
    E = 256
    q, k, v = (t.view(2, 19, 8, 32).transpose(1, 2) for t in qkv.split(E, dim=-1))
    ref = (torch.softmax(q @ k.transpose(-1, -2) / 32 ** 0.5, -1) @ v).transpose(1, 2).reshape(2, 19, E)
    assert rel_l2(o, ref) < 1e-13
    del sd


@pytest.mark.parametrize("kind,C", [("cal_si_snr", 2), ("pit", 2), ("pit", 3), ("tm", 1)])
def test_manual_loss_gradients_match_autograd(kind, C):
    B, T = 2, 1500
    _, a = synth.make_mixture(B, T, max(C, 2), seed=21)
    _, b = synth.make_mixture(B, T, max(C, 2), seed=22)
    a, b = a[:, :, :C].double(), b[:, :, :C].double()
    est = 0.6 * a + 0.4 * b.flip(-1)
    if kind == "cal_si_snr":
        src_t, est_t = a.transpose(0, 1), est.transpose(0, 1)            # [T,B,C]
        _, ds_ref, de_ref = BO.autograd_loss(kind, src_t, est_t)
        ds = torch.zeros_like(a)
        de = torch.zeros_like(a)
        for bi in range(B):
            for c in range(C):
                ds[bi, :, c], de[bi, :, c] = BO.manual_sb_pair_grad(a[bi, :, c], est[bi, :, c])
        assert rel_l2(ds.transpose(0, 1), ds_ref) < 1e-9
        assert rel_l2(de.transpose(0, 1), de_ref) < 1e-9
    elif kind == "pit":
        _, perms = O.pit_si_snr(est, a)
        _, ds_ref, de_ref = BO.autograd_loss(kind, est, a)                # training order: (estimate, targets)
        ds, de = BO.manual_pit_grad(est, a, perms)
        assert rel_l2(ds, ds_ref) < 1e-9
        assert rel_l2(de, de_ref) < 1e-9
    else:
        p, t = est[:, :, 0], a[:, :, 0]
        # torchmetrics takes eps = finfo(input dtype).eps: differentiate in float32 like the reference does
        _, dp_ref, dt_ref = BO.autograd_loss(kind, p.float(), t.float())
        dp, dt = BO.manual_tm_grad(p, t)
        assert rel_l2(dp, dp_ref) < 1e-5
        assert rel_l2(dt, dt_ref) < 1e-5


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="live reference not present")
def test_oracle_layer_gradients_match_live_reference():
    """Pin: autograd over the reference's own TransformerEncoderLayer == autograd over the oracle."""
    from oracle import run_reference
    run_reference._paths()
    from src.models import CSE_transformer as ref_mod
    layer = ref_mod.TransformerEncoderLayer(d_ffn=1024, nhead=8, d_model=256, dropout=0.0, normalize_before=True)
    layer = layer.double().eval()
    p = _layer_params(31)
    layer.load_state_dict(p, strict=True)
    g = torch.Generator().manual_seed(32)
    x = torch.randn(2, 23, 256, generator=g).double()
    dy = torch.randn(2, 23, 256, generator=g).double()
    xi = x.clone().requires_grad_(True)
    out = layer(xi)
    out = out[0] if isinstance(out, tuple) else out
    out.backward(dy)
    y, dx, grads = BO.autograd_layer(p, x, dy)
    assert rel_l2(y, out.detach()) < 1e-12
    assert rel_l2(dx, xi.grad) < 1e-11
    named = dict(layer.named_parameters())
    for k in BO.LAYER_KEYS:
        assert rel_l2(grads[k], named[k].grad) < 1e-11, k


def _oracle_training_grads(model_name, kind):
    from helpers import model_case
    sd, mix, src, ctx, se, meta = model_case(model_name)
    sd64 = {k: (v.double().requires_grad_(True) if "pos_enc" not in k else v) for k, v in sd.items()}
    out = O.sepformer_forward(sd64, mix.double(), ctx.double(), meta["variant"], meta["spk"])
    if kind == "tm_neg_sisnr":
        loss = -O.tm_si_snr(out[:, :, 0].float(), src[:, :, 0].float()).mean().double()
    else:
        est, pred = out
        pit, _ = O.pit_si_snr(est, src[:, :, : meta["spk"]].double())
        loss = pit.mean() + 0.1 * torch.logsumexp(pred, -1).mean()
    loss.backward()
    return loss.item(), {k: v.grad for k, v in sd64.items() if v.requires_grad}


@pytest.mark.parametrize("name", ["grad_context_2spk_b2_t3000", "grad_contsep_2spk_bce_b1_t2024"])
def test_oracle_gradients_match_reference_golden(name):
    """Pin of the gradient oracle: autograd over oracle/sepformer_oracle.py against the gradients the
    REFERENCE's own modules and loss objects produced (tests/golden/make_golden_grads.py)."""
    import numpy as np
    from cases import GRAD_CASES, GRAD_HEAD
    from helpers import GOLDEN
    model_name, kind = GRAD_CASES[name]
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        gold_loss, names = float(z["loss"]), [str(n) for n in z["names"]]
        norms, heads = torch.from_numpy(z["norms"]), torch.from_numpy(z["heads"])
    loss, grads = _oracle_training_grads(model_name, kind)
    assert abs(loss - gold_loss) < 1e-3, (loss, gold_loss)
    assert set(names) == set(grads)
    total = norms.norm().item()
    worst = 0.0
    for i, k in enumerate(names):
        g = grads[k] if grads[k] is not None else torch.zeros(1, dtype=torch.float64)
        # the fixture is the reference's fp32 arithmetic: compare at fp32-accumulation accuracy, relative
        # to the parameter's own gradient norm (floored for the parameters whose gradient is ~0)
        scale = max(norms[i].item(), 1e-6 * total)
        assert abs(g.norm().item() - norms[i].item()) < 3e-3 * scale, k
        flat = g.flatten()[:GRAD_HEAD]
        err = (flat - heads[i, : flat.numel()].double()).norm().item() / scale
        worst = max(worst, err)
        assert err < 3e-3, (k, err)
    print(f"{name}: oracle vs reference gradients, worst head error {worst:.2e} of the parameter's gradient norm")


def test_bf16_backward_design_within_reference_drift():
    """The planned tensor-core training step (bf16 operands and bf16 incoming gradients in every
    contraction, fp32 everywhere else — oracle/bf16_emulation.py) against the fp64 gradients: its drift must
    not exceed the drift of the reference's OWN bf16-autocast step recorded in the golden fixture."""
    import numpy as np
    from helpers import GOLDEN
    from oracle.bf16_emulation import emulate_bf16_gemms
    name, (model_name, kind) = "grad_context_2spk_b2_t3000", ("context_2spk_b2_t3000", "tm_neg_sisnr")
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        ref_drift, ref_loss16, gold_loss = float(z["bf16_ref_global_rel_l2"]), float(z["bf16_ref_loss"]), float(z["loss"])
    _, g_ref = _oracle_training_grads(model_name, kind)
    with emulate_bf16_gemms():
        loss16, g16 = _oracle_training_grads(model_name, kind)
    num = sum(((g16[k] - g_ref[k]) ** 2).sum() for k in g_ref if g_ref[k] is not None)
    den = sum((g_ref[k] ** 2).sum() for k in g_ref if g_ref[k] is not None)
    drift = (num / den).sqrt().item()
    print(f"bf16-GEMM emulation: loss {loss16:.4f} (fp32 {gold_loss:.4f}, reference autocast {ref_loss16:.4f}); "
          f"gradient drift {drift:.3f} vs reference autocast drift {ref_drift:.3f}")
    assert drift > 1e-4, "the emulation did not engage"
    assert drift < 1.25 * ref_drift
    assert abs(loss16 - gold_loss) < 1.25 * abs(ref_loss16 - gold_loss) + 0.05


def test_performance_mode_layer_backward_tolerance_is_reachable():
    """Numerical model of cse_layer_bwd_bf16 (bf16 wherever that composition stores or multiplies bf16 values)
    against the fp64 gradients: sets the bound of tests/test_backward_tc_gpu.py before it has run (one layer at
    bf16 moves its gradients by 2-5 %; 32 of them give the 0.12 whole-model drift of the test above)."""
    p = _layer_params(71)
    g = torch.Generator().manual_seed(72)
    x = torch.randn(3, 251, 256, generator=g).double()
    dy = torch.randn(3, 251, 256, generator=g).double()
    _, dx_ref, g_ref = BO.autograd_layer(p, x, dy)
    rnd = lambda t: t.to(torch.bfloat16).to(t.dtype)
    dx, grads = BO.manual_layer_bwd(p, x, dy, rnd=rnd)
    errs = {k: rel_l2(grads[k], g_ref[k]) for k in BO.LAYER_KEYS}
    errs["dx"] = rel_l2(dx, dx_ref)
    print("bf16 layer-backward model, rel-L2 per gradient:", {k: f"{v:.1e}" for k, v in errs.items()})
    assert 1e-4 < max(errs.values()) < 6e-2            # engaged, and inside the GPU test's 8e-2 bound


def test_tensor_core_linear_backward_dataflow():
    """Index algebra of cse_linear_bwd_tc (csrc/backward_tc.cu), replayed with the semantics of its building
    blocks — gemm_tc(A, W) = A W^T over K-contiguous operands, transposing casts zero-padded to Mpad — against the
    closed form.  (Numerics aside: this is about which matrix is transposed where.)"""
    M, N, K = 70, 256, 128
    g = torch.Generator().manual_seed(5)
    a, w, dc = (torch.randn(M, K, generator=g).double(), torch.randn(N, K, generator=g).double(),
                torch.randn(M, N, generator=g).double())
    da_ref, dw_ref, _ = BO.manual_linear_bwd(a, w, dc)
    gemm_tc = lambda A, W: A @ W.t()                       # C[M,N] = A[M,K] W[N,K]^T

    def transpose_cast(X, Mpad):                           # [M,n] -> [n,Mpad], zero padded
        out = torch.zeros(X.shape[1], Mpad, dtype=X.dtype)
        out[:, : X.shape[0]] = X.t()
        return out

    Mpad = (M + 63) // 64 * 64
    wt = w.t().contiguous()                                # launch_transpose(W, N, K) -> [K,N]
    da = gemm_tc(dc, wt)                                   # (M, N_gemm=K, K_gemm=N)
    dw = gemm_tc(transpose_cast(dc, Mpad), transpose_cast(a, Mpad))   # (M_gemm=N, N_gemm=K, K_gemm=Mpad)
    assert da.shape == (M, K) and dw.shape == (N, K)
    assert rel_l2(da, da_ref) < 1e-12 and rel_l2(dw, dw_ref) < 1e-12
