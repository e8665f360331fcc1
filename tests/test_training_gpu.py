"""GPU: the training path (cse_b200/training.py) — every non-transformer stage's backward kernel
against autograd over the CPU oracle, then whole-model gradients of the reference's training losses
(train_ContExt.py:366-367, train_ContSep.py:386-394) through the module API."""
import ctypes as C

import pytest
import torch

import cse_b200  # noqa: F401
from cse_b200 import _lib, losses, training
from helpers import model_case, rel_l2
from oracle import sepformer_oracle as O
from test_forward_gpu import build_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def _leaf(t, dev=None):
    t = t.to(dev) if dev else t.double()
    return t.detach().clone().requires_grad_(True)


def _cl(x):
    """reference [B,N,...] -> channels-last [B,...,N]"""
    return x.movedim(1, -1).contiguous()


# ------------------------------------------------------------------------------------------
# stages
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,T", [(2, 4003), (1, 16)])
def test_encoder_backward(B, T):
    w, mix = _rand(256, 1, 16, seed=1) / 4, _rand(B, T, seed=2) * 0.3
    wr = _leaf(w)
    ref = O.encoder({"encoder.conv1d.weight": wr}, mix.double())            # [B,256,L]
    dy = _rand(*ref.shape, seed=3)
    ref.backward(dy.double())
    wg = _leaf(w, DEV)
    out = training.EncoderFn.apply(mix.to(DEV), wg)
    assert rel_l2(out.detach().cpu(), _cl(ref.detach())) < 1e-6
    out.backward(_cl(dy).to(DEV))
    assert rel_l2(wg.grad.cpu(), wr.grad) < TOL


@pytest.mark.parametrize("B,rows,skip", [(2, 999, True), (3, 64, False), (1, 1, False)])
def test_groupnorm_forward_backward(B, rows, skip):
    x = _rand(B, rows, 256, seed=4) * 2 + 0.3
    g, b = 1 + 0.1 * _rand(256, seed=5), 0.1 * _rand(256, seed=6)
    sk = _rand(B, rows, 256, seed=7) if skip else None
    dy = _rand(B, rows, 256, seed=8)
    xr, gr, br = _leaf(x), _leaf(g), _leaf(b)
    skr = _leaf(sk) if skip else None
    ref = O.group_norm1(xr.transpose(1, 2), gr, br).transpose(1, 2)
    if skip:
        ref = ref + skr
    ref.backward(dy.double())
    xg, gg, bg = _leaf(x, DEV), _leaf(g, DEV), _leaf(b, DEV)
    skg = _leaf(sk, DEV) if skip else None
    out = training.GroupNormFn.apply(xg, gg, bg, skg)
    assert rel_l2(out.detach().cpu(), ref.detach()) < 1e-5
    out.backward(dy.to(DEV))
    assert rel_l2(xg.grad.cpu(), xr.grad) < TOL
    assert rel_l2(gg.grad.cpu(), gr.grad) < TOL
    assert rel_l2(bg.grad.cpu(), br.grad) < TOL
    if skip:
        assert torch.equal(skg.grad.cpu(), dy)


@pytest.mark.parametrize("B,L", [(2, 249), (1, 700)])
def test_segment_and_overlap_add_backward(B, L):
    x0 = _rand(B, L, 256, seed=9)
    S = _lib.path_shape(B, 8 * (L - 1) + 16, 0, 2).S
    # segmentation
    xr = _leaf(x0)
    seg, gap = O.pad_and_segment(xr.transpose(1, 2))                          # [B,N,K,S]
    dseg = _rand(*seg.shape, seed=10)
    seg.backward(dseg.double())
    xg = _leaf(x0, DEV)
    X = training.SegmentFn.apply(xg, S)                                       # [B,S,K,256]
    assert torch.equal(X.detach().cpu(), seg.detach().permute(0, 3, 2, 1).float())
    X.backward(dseg.permute(0, 3, 2, 1).contiguous().to(DEV))
    assert rel_l2(xg.grad.cpu(), xr.grad) < 1e-6
    # PReLU + overlap-add
    Xc = _rand(B, S, 250, 256, seed=11)
    a = torch.tensor([0.25])
    Xr, ar = _leaf(Xc), _leaf(a)
    xp = Xr.permute(0, 3, 2, 1)
    u = O.overlap_add(torch.where(xp >= 0, xp, ar * xp), gap)                 # [B,N,L]
    du = _rand(*u.shape, seed=12)
    u.backward(du.double())
    Xg, ag = _leaf(Xc, DEV), _leaf(a, DEV)
    U = training.PreluOverlapAddFn.apply(Xg, ag, L)
    assert rel_l2(U.detach().cpu(), _cl(u.detach())) < 1e-6
    U.backward(_cl(du).to(DEV))
    assert rel_l2(Xg.grad.cpu(), Xr.grad) < 1e-6
    assert rel_l2(ag.grad.cpu(), ar.grad) < TOL


@pytest.mark.parametrize("inter", [False, True])
@pytest.mark.parametrize("c", [0, 2])
def test_sequence_relayouts_are_adjoint_and_match_torch(inter, c):
    B, S = 2, 5
    X = _rand(B, S, 250, 256, seed=13)
    tok = _rand(B, c, 256, seed=14) if c else None
    pe = _rand(2500, 256, seed=15)
    nseq, n = (B * 250, S + c) if inter else (B * S, 250 + c)
    Xg = _leaf(X, DEV)
    tg = _leaf(tok, DEV) if c else None
    R = training.BuildSequencesFn.apply(Xg, tg, pe.to(DEV), inter)
    seqs = X.permute(0, 2, 1, 3).reshape(nseq, S, 256) if inter else X.reshape(nseq, 250, 256)
    if c:
        per_b = nseq // B
        seqs = torch.cat([tok.unsqueeze(1).expand(B, per_b, c, 256).reshape(nseq, c, 256), seqs], 1)
    ref = seqs + pe[:n]
    assert torch.equal(R.detach().cpu().view(nseq, n, 256), ref)
    dR = _rand(nseq * n, 256, seed=16)
    R.backward(dR.to(DEV))
    d3 = dR.view(nseq, n, 256)
    body = d3[:, c:]
    dX_ref = body.reshape(B, 250, S, 256).permute(0, 2, 1, 3) if inter else body.reshape(B, S, 250, 256)
    assert torch.equal(Xg.grad.cpu(), dX_ref.contiguous())
    if c:
        assert rel_l2(tg.grad.cpu(), d3[:, :c].reshape(B, nseq // B, c, 256).sum(1)) < 1e-6
    # and the way back
    Rg = _leaf(dR, DEV)
    Y, tok_sum = training.SequencesToChunksFn.apply(Rg, B, S, c, inter)
    assert torch.equal(Y.detach().cpu(), dX_ref.contiguous())
    w = _rand(B, S, 250, 256, seed=17)
    wt = _rand(B, c, 256, seed=18)
    ((Y * w.to(DEV)).sum() + (tok_sum * wt.to(DEV)).sum()).backward()
    back = w.permute(0, 2, 1, 3).reshape(nseq, S, 256) if inter else w.reshape(nseq, 250, 256)
    if c:
        back = torch.cat([wt.unsqueeze(1).expand(B, nseq // B, c, 256).reshape(nseq, c, 256), back], 1)
    assert torch.equal(Rg.grad.cpu().view(nseq, n, 256), back)


def test_gate_linear_and_context_map_backward():
    M = 333
    o, g, dy = _rand(M, 256, seed=19), _rand(M, 256, seed=20), _rand(M, 256, seed=21)
    orf, grf = _leaf(o), _leaf(g)
    (torch.tanh(orf) * torch.sigmoid(grf)).backward(dy.double())
    og, gg = _leaf(o, DEV), _leaf(g, DEV)
    training.GateFn.apply(og, gg).backward(dy.to(DEV))
    assert rel_l2(og.grad.cpu(), orf.grad) < TOL
    assert rel_l2(gg.grad.cpu(), grf.grad) < TOL
    # conv2d flavour: bias counted twice (overlap-add commuted in front of the 1x1 conv)
    W, b = _rand(512, 256, 1, 1, seed=22) / 16, _rand(512, seed=23)
    a = _rand(M, 256, seed=24)
    dy2 = _rand(M, 512, seed=25)
    Wr, br, ar = _leaf(W), _leaf(b), _leaf(a)
    (ar @ Wr[:, :, 0, 0].t() + 2 * br).backward(dy2.double())
    Wg, bg, ag = _leaf(W, DEV), _leaf(b, DEV), _leaf(a, DEV)
    y = training.LinearFn.apply(ag, Wg, bg, 2.0)
    y.backward(dy2.to(DEV))
    assert rel_l2(ag.grad.cpu(), ar.grad) < TOL
    assert rel_l2(Wg.grad.cpu(), Wr.grad) < TOL and Wg.grad.shape == W.shape
    assert rel_l2(bg.grad.cpu(), br.grad) < TOL
    # context mapper
    x, Wm, bm = _rand(3, 4096, seed=26), _rand(256, 4096, seed=27) / 64, _rand(256, seed=28)
    dt = _rand(3, 256, seed=29)
    xr, Wmr, bmr = _leaf(x), _leaf(Wm), _leaf(bm)
    (xr @ Wmr.t() + bmr).backward(dt.double())
    xg, Wmg, bmg = _leaf(x, DEV), _leaf(Wm, DEV), _leaf(bm, DEV)
    training.ContextMapFn.apply(xg, Wmg, bmg).backward(dt.to(DEV))
    assert rel_l2(Wmg.grad.cpu(), Wmr.grad) < TOL
    assert rel_l2(bmg.grad.cpu(), bmr.grad) < TOL
    assert rel_l2(xg.grad.cpu(), xr.grad) < TOL


@pytest.mark.parametrize("B,L,T,n_masks", [(2, 249, 2000, 2), (1, 124, 1010, 1), (1, 30, 240, 3)])
def test_mask_decode_backward(B, L, T, n_masks):
    """T > T_est exercises the zero-pad branch (ContSep.py:92-93), T < T_est the trim branch (:95)."""
    mp = _rand(B * L * n_masks, 256, seed=30)
    E = _rand(B, L, 256, seed=31).abs()
    w = _rand(256, 1, 16, seed=32) / 4
    d_est = _rand(B, T, n_masks, seed=33)
    mpr, Er, wr = _leaf(mp), _leaf(E), _leaf(w)
    mask = torch.relu(mpr).view(B, L, n_masks, 256)
    ests = [O.decoder({"decoder.weight": wr}, (mask[:, :, s] * Er).transpose(1, 2)) for s in range(n_masks)]
    est = O.fix_length(torch.stack(ests, -1), T)
    est.backward(d_est.double())
    mpg, Eg, wg = _leaf(mp, DEV), _leaf(E, DEV), _leaf(w, DEV)
    out = training.MaskDecodeFn.apply(mpg, Eg, wg, T, n_masks)
    assert rel_l2(out.detach().cpu(), est.detach()) < 1e-5
    out.backward(d_est.to(DEV))
    assert rel_l2(mpg.grad.cpu(), mpr.grad) < TOL
    assert rel_l2(Eg.grad.cpu(), Er.grad) < TOL
    assert rel_l2(wg.grad.cpu(), wr.grad) < TOL


# ------------------------------------------------------------------------------------------
# whole model: gradients of the reference's training losses
# ------------------------------------------------------------------------------------------
def _reference_grads(sd, mix, src, ctx, se, meta, loss_fn):
    sd64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "pos_enc" not in k else v)
            for k, v in sd.items()}
    out = O.sepformer_forward(sd64, mix.double(), None if ctx is None else ctx.double(), meta["variant"],
                              meta["spk"], None if se is None else se.double(), meta["cue"] or "joint")
    loss = loss_fn(out, src.double(), O)
    loss.backward()
    return loss.item(), {k: v.grad for k, v in sd64.items() if getattr(v, "grad", None) is not None}


def _ours(name, loss_fn):
    sd, mix, src, ctx, se, meta = model_case(name)
    m = build_model(meta)
    m.load_state_dict(sd)
    m = m.to(DEV).train()
    m.precision = "fp32"
    args = [mix.to(DEV)]
    if meta["variant"] != "sepformer":
        args.append(ctx.to(DEV))
    kw = {}
    if meta["variant"] == "hcontext":
        m.eval()                                   # the train-mode cue choice is random (ContExt.py:98-104)
        args.append(se.to(DEV))
        kw["cue"] = meta["cue"]
    out = m(*args, **kw)
    loss = loss_fn(out, src.to(DEV), None)
    loss.backward()
    torch.cuda.synchronize()
    ref_loss, ref = _reference_grads(sd, mix, src, ctx, se, meta, loss_fn)
    got = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    return loss.item(), got, ref_loss, ref


def _loss_context(out, src, oracle):
    est = out[:, :, 0]                                                        # train_ContExt.py:366-367
    if oracle is not None:
        return -oracle.tm_si_snr(est.float(), src[:, :, 0].float()).mean().double()
    return -losses.ScaleInvariantSignalNoiseRatio()(est, src[:, :, 0])


def _loss_contsep(out, src, oracle):
    est, pred = out                                                           # train_ContSep.py:391-394
    if oracle is not None:
        pit, _ = oracle.pit_si_snr(est, src[:, :, : est.shape[-1]])
        return pit.mean() + 0.1 * torch.logsumexp(pred, -1).mean()
    pit = losses.get_si_snr_with_pitwrapper(est, src[:, :, : est.shape[-1]].contiguous())
    return pit.mean() + 0.1 * torch.logsumexp(pred, -1).mean()


def _loss_sepformer(out, src, oracle):
    if oracle is not None:
        pit, _ = oracle.pit_si_snr(out, src[:, :, : out.shape[-1]])
        return pit.mean()
    return losses.get_si_snr_with_pitwrapper(out, src[:, :, : out.shape[-1]].contiguous()).mean()


@pytest.mark.parametrize("name,loss_fn", [
    ("context_2spk_b2_t3000", _loss_context),           # BASELINE configs[2] flavour: ContExt, B=2
    ("contsep_2spk_bce_b1_t2024", _loss_contsep),       # pred_head / context_selector branch, T % 8 != 0
    ("sepformer_3spk_b2_t1000", _loss_sepformer),       # c = 0, three masks
    ("hcontext_2spk_voice_b1_t1999", _loss_context),    # c = 2, gradient reaches se_embedding through ctx
])
def test_whole_model_gradients(name, loss_fn):
    loss, got, ref_loss, ref = _ours(name, loss_fn)
    assert abs(loss - ref_loss) < 2e-3 * max(1.0, abs(ref_loss)), (loss, ref_loss)
    missing = [k for k in ref if k not in got and ref[k].abs().max() > 0]
    assert not missing, missing
    num = sum(((got[k].cpu().double() - ref[k]) ** 2).sum() for k in ref if k in got)
    den = sum((ref[k] ** 2).sum() for k in ref if k in got)
    worst = max((rel_l2(got[k].cpu(), ref[k]), k) for k in ref if k in got and ref[k].norm() > 1e-9)
    print(f"{name}: loss {loss:.5f} (ref {ref_loss:.5f}), global grad rel-L2 {(num / den).sqrt():.2e}, worst {worst}")
    assert (num / den).sqrt().item() < 1e-3
    assert worst[0] < 1e-2, worst


def test_training_step_updates_parameters_like_the_oracle():
    """One SGD step on the ContExt loss moves the loss the same way on both sides."""
    sd, mix, src, ctx, se, meta = model_case("context_2spk_c3_b1_t2000")
    m = build_model(meta)
    m.load_state_dict(sd)
    m = m.to(DEV).train()
    m.precision = "fp32"
    opt = torch.optim.SGD(m.parameters(), lr=1e-3)
    vals = []
    for _ in range(2):
        opt.zero_grad()
        loss = _loss_context(m(mix.to(DEV), ctx.to(DEV)), src.to(DEV), None)
        loss.backward()
        opt.step()
        vals.append(loss.item())
    sd64 = {k: (v.double().requires_grad_(True) if "pos_enc" not in k else v) for k, v in sd.items()}
    ref_vals = []
    for _ in range(2):
        out = O.sepformer_forward(sd64, mix.double(), ctx.double(), "context", 2)
        loss = _loss_context(out, src.double(), O)
        grads = torch.autograd.grad(loss, [v for v in sd64.values() if v.requires_grad], allow_unused=True)
        with torch.no_grad():
            for v, g in zip([v for v in sd64.values() if v.requires_grad], grads):
                if g is not None:
                    v -= 1e-3 * g
        ref_vals.append(loss.item())
    assert abs(vals[0] - ref_vals[0]) < 2e-3 and abs(vals[1] - ref_vals[1]) < 5e-3, (vals, ref_vals)


@pytest.mark.parametrize("name,loss_fn,fixture", [
    ("context_2spk_b2_t3000", _loss_context, "grad_context_2spk_b2_t3000"),
    ("contsep_2spk_bce_b1_t2024", _loss_contsep, "grad_contsep_2spk_bce_b1_t2024"),
])
@pytest.mark.parametrize("amp_dtype", [torch.bfloat16, torch.float16])
def test_autocast_training_step_carries_a_graph_and_stays_within_reference_drift(name, loss_fn, fixture, amp_dtype):
    """The reference's actual training mode (train_ContSep.py:383-400, README.md:142: forward + loss under
    torch.autocast, scaler.scale(loss).backward()): the drop-in must return outputs WITH an autograd graph, and
    the gradients of the bf16 tensor-core path may drift from the fp32 gradients by no more than the reference's
    own autocast step does on the same fixture (`bf16_ref_global_rel_l2`, make_golden_grads.py)."""
    from helpers import load_golden
    sd, mix, src, ctx, se, meta = model_case(name)
    m = build_model(meta)
    m.load_state_dict(sd)
    m = m.to(DEV).train()
    scaler = torch.amp.GradScaler("cuda", enabled=(amp_dtype == torch.float16))
    with torch.autocast("cuda", dtype=amp_dtype):
        out = m(mix.to(DEV), ctx.to(DEV))
        first = out[0] if isinstance(out, tuple) else out
        assert first.requires_grad and first.grad_fn is not None, "autocast training output carries no graph"
        loss = loss_fn(out, src.to(DEV), None)
    scaler.scale(loss).backward()
    torch.cuda.synchronize()
    inv = 1.0 / scaler.get_scale() if amp_dtype == torch.float16 else 1.0
    ref_loss, ref = _reference_grads(sd, mix, src, ctx, se, meta, loss_fn)
    got = {k: p.grad * inv for k, p in m.named_parameters() if p.grad is not None}
    missing = [k for k in ref if k not in got and ref[k].abs().max() > 0]
    assert not missing, missing
    num = sum(((got[k].cpu().double() - ref[k]) ** 2).sum() for k in ref if k in got)
    den = sum((ref[k] ** 2).sum() for k in ref if k in got)
    drift = (num / den).sqrt().item()
    fix = load_golden(fixture)
    bar = float(fix["bf16_ref_global_rel_l2"])
    print(f"\n[autocast {amp_dtype} {name}] loss {loss.item():.4f} (fp64 oracle {ref_loss:.4f}, reference bf16 "
          f"{float(fix['bf16_ref_loss']):.4f}); global gradient drift {drift:.3e} vs the reference's own autocast drift {bar:.3e}")
    assert all(torch.isfinite(g).all() for g in got.values())
    assert drift <= bar
    assert abs(loss.item() - ref_loss) <= max(0.05, 1.5 * abs(float(fix["bf16_ref_loss"]) - float(fix["loss"])))


def test_graphed_training_step_matches_eager_steps():
    """runtime.GraphedStep: the autocast step (forward, -SI-SNR, backward, clip + fused AdamW) replayed as one CUDA
    graph follows the same trajectory as eagerly launched steps (train_ContExt.py:365-389).  Same kernels, same
    inputs; the only run-to-run difference is the order of the split-K reduce-adds of the weight gradients."""
    from cse_b200.optim import AdamW
    from cse_b200.runtime import GraphedStep
    sd, mix, src, ctx, se, meta = model_case("context_2spk_b2_t3000")
    mix_d, ctx_d, tgt_d = mix.to(DEV), ctx.to(DEV), src[:, :, 0].contiguous().to(DEV)
    sisnr = losses.ScaleInvariantSignalNoiseRatio()

    def make():
        m = build_model(meta)
        m.load_state_dict(sd)
        m = m.to(DEV).train()
        opt = AdamW(m.parameters(), lr=1e-3, amsgrad=True)

        def step(mx, cx, tg):
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = -sisnr(m(mx, cx)[:, :, 0], tg)
            loss.backward()
            opt.step(max_norm=5.0)
            return loss
        return m, opt, step

    m_e, opt_e, step_e = make()
    eager_losses = [step_e(mix_d, ctx_d, tgt_d).item() for _ in range(6)]

    m_g, opt_g, step_g = make()
    graphed = GraphedStep(step_g, mix_d.clone(), ctx_d.clone(), tgt_d.clone(), warmup=3)   # three eager steps, then the capture
    graph_losses = [graphed(mix_d, ctx_d, tgt_d).item() for _ in range(3)]                 # steps 4-6
    torch.cuda.synchronize()
    assert opt_g.steps_applied() == 6 and opt_e.steps_applied() == 6
    num = sum(((pg.detach() - pe.detach()).double() ** 2).sum() for pg, pe in zip(m_g.parameters(), m_e.parameters()))
    den = sum((pe.detach().double() ** 2).sum() for pe in m_e.parameters())
    moved = sum(((pe.detach().cpu() - sd[k]).double() ** 2).sum() for k, pe in m_e.named_parameters())
    drift = (num / den).sqrt().item()
    print(f"\n[graphed step] losses eager {eager_losses} graph {graph_losses}; parameter rel-L2 graph vs eager {drift:.2e}, "
          f"eager vs initial {(moved / den).sqrt().item():.2e}")
    # Adam's update is lr * m / sqrt(v): where a gradient is near zero its SIGN (which the reduce-add order can flip)
    # moves the parameter by the full lr, so the two trajectories agree to a fraction of the distance travelled, not
    # to rounding error
    assert all(abs(a - b) <= 5e-2 * max(1.0, abs(b)) for a, b in zip(graph_losses, eager_losses[3:])), (graph_losses, eager_losses)
    assert eager_losses[-1] < eager_losses[0] and graph_losses[-1] < eager_losses[0]   # the steps do train
    assert drift < 0.5 * (moved / den).sqrt().item()
    # an eager step of the same optimiser after the capture still works (its own staging copies and device table)
    before = opt_g.steps_applied()
    step_g(mix_d, ctx_d, tgt_d)
    graphed(mix_d, ctx_d, tgt_d)
    torch.cuda.synchronize()
    assert opt_g.steps_applied() == before + 2
    assert all(torch.isfinite(p).all() for p in m_g.parameters())
