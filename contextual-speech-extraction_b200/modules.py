"""Host-side mirror of the reference's module API for the Sepformer hot path.

Same class names, constructor arguments, parameter names/shapes and return conventions as
`src/models/{ContSep,ContExt,sepformer,CSE_transformer}.py` and the speechbrain leaves they
import, so `load_state_dict(ckpt['state_dict'])` and the train/test scripts work
unchanged.  The modules own ordinary fp32 `nn.Parameter`s; every forward() launches the
hand-written sm_100a kernels through the C ABI (`include/cse_b200.h`).  There is no PyTorch
compute fallback: without the built library or on a CPU tensor the call raises.

The kernels are specialised to the constants hard-coded in the reference constructors
(ContSep.py:10-40): N=256, encoder k=16/s=8, K=250, 8 heads, d_ffn=1024, 8 layers, pre-norm,
ReLU, dropout 0, norm='ln', no linear after intra/inter, skip around intra.
"""
import copy
import ctypes as C
import math
import warnings

import torch
import torch.nn as nn

from . import _lib
from ._lib import BF16, FP32
from .runtime import WORKSPACE, ParamTable, current_stream, resolve_precision
from .shapes import CHUNK, D_FFN, ENC_K, ENC_S, N_CH, N_HEAD, N_LAYER, PE_MAX

EPS = 1e-8


def _unsupported(what):
    raise NotImplementedError(
        f"{what}: the B200 kernels are specialised to the reference configuration "
        "(N=256, kernel 16/stride 8, K=250, 8 heads, d_ffn=1024, 8 pre-norm ReLU layers, norm='ln')")


def _act_dtype(precision):
    return torch.bfloat16 if precision == BF16 else torch.float32


def _check_cuda(t, name):
    if not t.is_cuda:
        raise _lib.CseError(f"{name} is on {t.device}: the CUDA path has no CPU fallback")


_warned_grad = [False]


def _warn_no_grad(module):
    if torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters()) and not _warned_grad[0]:
        _warned_grad[0] = True
        warnings.warn("cse_b200: the bf16 / autocast backward kernels are not built yet — these outputs carry no "
                      "autograd graph; run the training step in fp32 (precision='fp32', no autocast)")


def select_norm(norm, dim, shape, eps=1e-8):
    """speechbrain select_norm: 'ln' -> GroupNorm(1, dim, eps) (ContSep.py:164,423-424)."""
    if norm == "ln":
        return nn.GroupNorm(1, dim, eps=eps)
    _unsupported(f"norm={norm!r}")


# ----------------------------------------------------------------------------------------------
# speechbrain leaves
# ----------------------------------------------------------------------------------------------
class Encoder(nn.Module):
    """speechbrain Encoder (ContSep.py:10,69): relu(Conv1d(1,256,16,stride 8,bias=False)).
    forward: [B,T] -> [B,256,L] (a transposed view of the channels-last kernel output)."""

    def __init__(self, kernel_size=2, out_channels=64, in_channels=1):
        super().__init__()
        self.conv1d = nn.Conv1d(in_channels=in_channels, out_channels=out_channels,
                                kernel_size=kernel_size, stride=kernel_size // 2, groups=1, bias=False)
        self.in_channels = in_channels

    def forward(self, x, precision=None):
        w = self.conv1d.weight
        if tuple(w.shape) != (N_CH, 1, ENC_K):
            _unsupported(f"Encoder weight {tuple(w.shape)}")
        if x.dim() != 2:
            raise RuntimeError(f"Encoder expects [B,T], got {tuple(x.shape)}")
        _check_cuda(x, "mix")
        prec = resolve_precision(precision)
        x = x.contiguous().float()
        B, T = x.shape
        sh = _lib.path_shape(B, T, 0, 1)
        out = torch.empty(B, sh.L, N_CH, dtype=_act_dtype(prec), device=x.device)
        part = torch.empty(B, (sh.L + 63) // 64, 2, dtype=torch.float32, device=x.device)
        n_parts = C.c_int(0)
        _lib.call("cse_encoder_fwd", _lib.ptr(x), _lib.ptr(w), B, T, prec, _lib.ptr(out), _lib.ptr(part),
                  C.byref(n_parts), C.c_void_p(current_stream(x.device)))
        return out.transpose(1, 2)


class Decoder(nn.ConvTranspose1d):
    """speechbrain Decoder (ContSep.py:40,84): ConvTranspose1d(256,1,16,stride=8,bias=False) whose
    forward takes [B,N,L] and returns [B,T_est]."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    def forward(self, x, precision=None):
        if x.dim() not in [2, 3]:
            raise RuntimeError("{} accept 3/4D tensor as input".format(type(self).__name__))
        if tuple(self.weight.shape) != (N_CH, 1, ENC_K) or self.bias is not None or self.stride != (ENC_S,):
            _unsupported(f"Decoder weight {tuple(self.weight.shape)}")
        if x.dim() == 2:
            x = x.unsqueeze(0)
        _check_cuda(x, "decoder input")
        B, N, L = x.shape
        rows = x.transpose(1, 2).contiguous().float()                # channels-last [B,L,N]
        T_est = ENC_S * (L - 1) + ENC_K
        frames = torch.empty(B * L, ENC_K, dtype=torch.float32, device=x.device)
        est = torch.empty(B, T_est, 1, dtype=torch.float32, device=x.device)
        _lib.call("cse_mask_decode", _lib.ptr(rows), None, _lib.ptr(self.weight), B, L, T_est, 1, FP32,
                  _lib.ptr(frames), _lib.ptr(est), C.c_void_p(current_stream(x.device)))
        return est.squeeze(-1)


class PositionalEncoding(nn.Module):
    """speechbrain PositionalEncoding (CSE_transformer.py:88): persistent sinusoid buffer `pe`."""

    def __init__(self, input_size, max_len=PE_MAX):
        super().__init__()
        if input_size % 2 != 0:
            raise ValueError(f"Cannot use sin/cos positional encoding with odd channels (got channels={input_size})")
        self.max_len = max_len
        pe = torch.zeros(max_len, input_size, requires_grad=False)
        pos = torch.arange(0, max_len).unsqueeze(1).float()
        freq = torch.exp(torch.arange(0, input_size, 2).float() * -(math.log(10000.0) / input_size))
        pe[:, 0::2] = torch.sin(pos * freq)
        pe[:, 1::2] = torch.cos(pos * freq)
        self.register_buffer("pe", pe.unsqueeze(0))

    def forward(self, x):
        return self.pe[:, : x.size(1)].clone().detach()


class LayerNorm(nn.Module):
    """sb.nnet.normalization.LayerNorm: parameter container `.norm` (CSE_transformer.py:197,358)."""

    def __init__(self, input_size=None, input_shape=None, eps=1e-05, elementwise_affine=True):
        super().__init__()
        self.eps = eps
        self.norm = nn.LayerNorm(input_size, eps=eps, elementwise_affine=elementwise_affine)


class PositionalwiseFeedForward(nn.Module):
    """sb.nnet.attention.PositionalwiseFeedForward: parameter container `.ffn.{0,3}`."""

    def __init__(self, d_ffn, input_shape=None, input_size=None, dropout=0.0, activation=nn.ReLU):
        super().__init__()
        self.ffn = nn.Sequential(nn.Linear(input_size, d_ffn), activation(), nn.Dropout(dropout),
                                 nn.Linear(d_ffn, input_size))


# ----------------------------------------------------------------------------------------------
# CSE_transformer.py mirror
# ----------------------------------------------------------------------------------------------
class MultiheadAttention(nn.Module):
    """CSE_transformer.py:424-477: parameter container `.att` = nn.MultiheadAttention."""

    def __init__(self, nhead, d_model, dropout=0.0, bias=True, add_bias_kv=False, add_zero_attn=False,
                 kdim=None, vdim=None):
        super().__init__()
        self.att = nn.MultiheadAttention(embed_dim=d_model, num_heads=nhead, dropout=dropout, bias=bias,
                                         add_bias_kv=add_bias_kv, add_zero_attn=add_zero_attn,
                                         kdim=kdim, vdim=vdim)


class TransformerEncoderLayer(nn.Module):
    """CSE_transformer.py:253-364 (regularMHA + regularFFN, pre-norm)."""

    def __init__(self, d_ffn, nhead, d_model, kdim=None, vdim=None, dropout=0.0, activation=nn.ReLU,
                 normalize_before=False, attention_type="regularMHA", ffn_type="regularFFN",
                 ffn_cnn_kernel_size_list=[3, 3], causal=False):
        super().__init__()
        if attention_type != "regularMHA" or ffn_type != "regularFFN" or causal:
            _unsupported(f"attention_type={attention_type!r}, ffn_type={ffn_type!r}, causal={causal}")
        self.self_att = MultiheadAttention(nhead=nhead, d_model=d_model, dropout=dropout, kdim=kdim, vdim=vdim)
        self.pos_ffn = PositionalwiseFeedForward(d_ffn=d_ffn, input_size=d_model, dropout=dropout,
                                                 activation=activation)
        self.norm1 = LayerNorm(d_model, eps=1e-6)
        self.norm2 = LayerNorm(d_model, eps=1e-6)
        self.dropout1 = torch.nn.Dropout(dropout)
        self.dropout2 = torch.nn.Dropout(dropout)
        self.normalize_before = normalize_before
        self.pos_ffn_type = ffn_type


class TransformerEncoder(nn.Module):
    """CSE_transformer.py:109-199: `layers` + final `norm` (eps 1e-6)."""

    def __init__(self, num_layers, nhead, d_ffn, input_shape=None, d_model=None, kdim=None, vdim=None,
                 dropout=0.0, activation=nn.ReLU, normalize_before=False, causal=False,
                 layerdrop_prob=0.0, attention_type="regularMHA", ffn_type="regularFFN",
                 ffn_cnn_kernel_size_list=[3, 3]):
        super().__init__()
        self.layers = torch.nn.ModuleList([
            TransformerEncoderLayer(d_ffn=d_ffn, nhead=nhead, d_model=d_model, kdim=kdim, vdim=vdim,
                                    dropout=dropout, activation=activation,
                                    normalize_before=normalize_before, causal=causal,
                                    attention_type=attention_type, ffn_type=ffn_type,
                                    ffn_cnn_kernel_size_list=ffn_cnn_kernel_size_list)
            for _ in range(num_layers)])
        self.norm = LayerNorm(d_model, eps=1e-6)
        self.layerdrop_prob = layerdrop_prob


class SBTransformerBlock_CSE(nn.Module):
    """CSE_transformer.py:11-106.  forward: [B',n,256] -> [B',n,256] = LN(8 layers(x + pe[:n]))."""

    def __init__(self, num_layers, d_model, nhead, d_ffn=2048, input_shape=None, kdim=None, vdim=None,
                 dropout=0.1, activation="relu", use_positional_encoding=False, norm_before=False,
                 attention_type="regularMHA"):
        super().__init__()
        self.use_positional_encoding = use_positional_encoding
        if activation == "relu":
            act = nn.ReLU
        elif activation == "gelu":
            act = nn.GELU
        else:
            raise ValueError("unknown activation")
        self._cfg = dict(num_layers=num_layers, d_model=d_model, nhead=nhead, d_ffn=d_ffn, dropout=dropout,
                         activation=activation, norm_before=norm_before, pe=use_positional_encoding)
        self.mdl = TransformerEncoder(num_layers=num_layers, nhead=nhead, d_ffn=d_ffn, input_shape=input_shape,
                                      d_model=d_model, kdim=kdim, vdim=vdim, dropout=dropout, activation=act,
                                      normalize_before=norm_before, attention_type=attention_type)
        if use_positional_encoding:
            self.pos_enc = PositionalEncoding(input_size=d_model)
        self._w16 = {}

    def _check_supported(self):
        c = self._cfg
        ok = (c["num_layers"] == N_LAYER and c["d_model"] == N_CH and c["nhead"] == N_HEAD
              and c["d_ffn"] == D_FFN and c["activation"] == "relu" and c["norm_before"] and c["pe"]
              and (c["dropout"] == 0 or not self.training))
        if not ok:
            _unsupported(f"SBTransformerBlock_CSE{c}")

    def _weight(self, w, prec):
        if prec == FP32:
            return w
        key = id(w)
        ent = self._w16.get(key)
        if ent is None or ent[0] != (w.data_ptr(), w._version):
            ent = ((w.data_ptr(), w._version), w.detach().to(torch.bfloat16).contiguous())
            self._w16[key] = ent
        return ent[1]

    def forward(self, x, precision=None):
        """Stand-alone block through the per-stage C-ABI entry points (the fused model path runs the
        same kernels from cse_forward)."""
        self._check_supported()
        _check_cuda(x, "x")
        prec = resolve_precision(precision)
        Bp, n, E = x.shape
        if E != N_CH:
            _unsupported(f"d_model={E}")
        dev = x.device
        st = C.c_void_p(current_stream(dev))
        R = (x.float() + self.pos_enc.pe[:, :n]).contiguous()          # residual stream, fp32
        M = Bp * n
        self._run_layers(R, Bp, n, prec)
        out = torch.empty(Bp, n, N_CH, dtype=torch.float32, device=dev)
        _lib.call("cse_layernorm_fwd", _lib.ptr(R), _lib.ptr(self.mdl.norm.norm.weight),
                  _lib.ptr(self.mdl.norm.norm.bias), M, 1e-6, FP32, _lib.ptr(out), st)
        return out

    def _run_layers(self, R, Bp, n, prec):
        """The 8 pre-norm layers, in place on the fp32 residual stream R [Bp*n, 256] (PE already
        added; the stack's final LayerNorm is applied by the caller / the stack tail)."""
        dev, adt = R.device, _act_dtype(prec)
        st = C.c_void_p(current_stream(dev))
        M = Bp * n
        H = torch.empty(M, N_CH, dtype=adt, device=dev)
        QKV = torch.empty(M, 3 * N_CH, dtype=adt, device=dev)
        AO = torch.empty(M, N_CH, dtype=adt, device=dev)
        F1 = torch.empty(M, D_FFN, dtype=adt, device=dev)
        for layer in self.mdl.layers:
            att, ffn = layer.self_att.att, layer.pos_ffn.ffn
            _lib.call("cse_layernorm_fwd", _lib.ptr(R), _lib.ptr(layer.norm1.norm.weight),
                      _lib.ptr(layer.norm1.norm.bias), M, 1e-6, prec, _lib.ptr(H), st)
            _lib.call("cse_linear", _lib.ptr(H), N_CH, _lib.ptr(self._weight(att.in_proj_weight, prec)),
                      _lib.ptr(att.in_proj_bias), 1.0, None, _lib.ptr(QKV), 3 * N_CH, M, 3 * N_CH, N_CH, 0, 0,
                      prec, st)
            _lib.call("cse_attention_fwd", _lib.ptr(QKV), Bp, n, prec, _lib.ptr(AO), st)
            _lib.call("cse_linear", _lib.ptr(AO), N_CH, _lib.ptr(self._weight(att.out_proj.weight, prec)),
                      _lib.ptr(att.out_proj.bias), 1.0, _lib.ptr(R), _lib.ptr(R), N_CH, M, N_CH, N_CH, 0, 1,
                      prec, st)
            _lib.call("cse_layernorm_fwd", _lib.ptr(R), _lib.ptr(layer.norm2.norm.weight),
                      _lib.ptr(layer.norm2.norm.bias), M, 1e-6, prec, _lib.ptr(H), st)
            _lib.call("cse_linear", _lib.ptr(H), N_CH, _lib.ptr(self._weight(ffn[0].weight, prec)),
                      _lib.ptr(ffn[0].bias), 1.0, None, _lib.ptr(F1), D_FFN, M, D_FFN, N_CH, 1, 0, prec, st)
            _lib.call("cse_linear", _lib.ptr(F1), D_FFN, _lib.ptr(self._weight(ffn[3].weight, prec)),
                      _lib.ptr(ffn[3].bias), 1.0, _lib.ptr(R), _lib.ptr(R), N_CH, M, N_CH, D_FFN, 0, 1, prec, st)


# ----------------------------------------------------------------------------------------------
# dual-path model
# ----------------------------------------------------------------------------------------------
class Dual_Computation_Block_CSE(nn.Module):
    """ContSep.py:372-533 / ContExt.py:398-557: parameter container for one dual-path block
    (intra_mdl, inter_mdl, intra_norm, inter_norm, optional context mappers)."""

    def __init__(self, intra_mdl, inter_mdl, out_channels, norm="ln", skip_around_intra=True,
                 linear_layer_after_inter_intra=True, llm_dim=4096):
        super().__init__()
        if linear_layer_after_inter_intra or not skip_around_intra or norm != "ln":
            _unsupported("Dual_Computation_Block with linear_layer_after_inter_intra / no skip / norm != 'ln'")
        self.intra_mdl = intra_mdl
        self.inter_mdl = inter_mdl
        self.skip_around_intra = skip_around_intra
        self.linear_layer_after_inter_intra = linear_layer_after_inter_intra
        self.llm_dim = llm_dim
        self.out_channels = out_channels
        self.norm = norm
        self.intra_norm = select_norm(norm, out_channels, 4)
        self.inter_norm = select_norm(norm, out_channels, 4)
        self.intra_context_mapper = None
        self.inter_context_mapper = None

    def add_ctx(self):
        if self.llm_dim != _lib.CTX:
            _unsupported(f"ctx_dim={self.llm_dim} (the context-mapper kernel is specialised to {_lib.CTX})")
        self.intra_context_mapper = nn.Linear(self.llm_dim, self.out_channels)
        self.inter_context_mapper = nn.Linear(self.llm_dim, self.out_channels)

    returns_pred_head = True      # ContSep flavour (ContSep.py:533); ContExt's returns `out` only

    def forward(self, x, ctx=None, precision=None):
        """ContSep.py:453-533 on its own, through the per-stage C-ABI entry points:
        x [B,N,K,S], ctx [B,c,4096] | None -> out [B,N,K,S] (, pred_head [B,N])."""
        self.intra_mdl._check_supported()
        self.inter_mdl._check_supported()
        _check_cuda(x, "x")
        B, N, K, S = x.shape
        if N != N_CH or K != CHUNK:
            _unsupported(f"Dual_Computation_Block input [B,{N},{K},S]")
        prec = resolve_precision(precision)
        dev = x.device
        st = C.c_void_p(current_stream(dev))
        X = x.permute(0, 3, 2, 1).float().contiguous()                 # [B,S,K,N] channels-last
        c = 0 if ctx is None else ctx.size(1)
        tok = [None, None]
        if c:
            ctx_d = ctx.contiguous().float()
            for i, mp in enumerate((self.intra_context_mapper, self.inter_context_mapper)):
                tok[i] = torch.empty(B, c, N_CH, dtype=torch.float32, device=dev)
                _lib.call("cse_context_map", _lib.ptr(ctx_d), _lib.ptr(mp.weight), _lib.ptr(mp.bias), B * c,
                          ctx_d.size(2), _lib.ptr(tok[i]), st)
        part = torch.empty(B, 64, 2, dtype=torch.float32, device=dev)
        stat = torch.empty(B, 2, dtype=torch.float32, device=dev)

        def stack(mdl, src, inter, norm, skip):
            n = (S if inter else K) + c
            nseq = B * (K if inter else S)
            R = torch.empty(nseq * n, N_CH, dtype=torch.float32, device=dev)
            _lib.call("cse_build_sequences", _lib.ptr(src), _lib.ptr(tok[inter]), _lib.ptr(mdl.pos_enc.pe), B, S, c,
                      inter, _lib.ptr(R), st)
            mdl._run_layers(R, nseq, n, prec)
            out = torch.empty(B, S, K, N_CH, dtype=torch.float32, device=dev)
            _lib.call("cse_stack_finish", _lib.ptr(R), _lib.ptr(mdl.mdl.norm.norm.weight),
                      _lib.ptr(mdl.mdl.norm.norm.bias), _lib.ptr(norm.weight), _lib.ptr(norm.bias), _lib.ptr(skip),
                      B, S, c, inter, _lib.ptr(out), _lib.ptr(part), _lib.ptr(stat), st)
            return R, out

        _, intra = stack(self.intra_mdl, X, 0, self.intra_norm, X)               # + x (skip around intra)
        R_inter, out = stack(self.inter_mdl, intra, 1, self.inter_norm, intra)   # + intra
        out = out.permute(0, 3, 2, 1)                                             # [B,N,K,S] view
        if not self.returns_pred_head:
            return out
        pred = torch.empty(B, N_CH, dtype=torch.float32, device=dev)
        _lib.call("cse_pred_head", _lib.ptr(R_inter), _lib.ptr(self.inter_mdl.mdl.norm.norm.weight),
                  _lib.ptr(self.inter_mdl.mdl.norm.norm.bias), B, S, c, _lib.ptr(pred), st)
        return out, pred


class Dual_Path_Model_CSE(nn.Module):
    """ContSep.py:103-370.  forward(x [B,N,L], ctx [B,c,4096] | None) ->
    (mask [spk,B,N,L], pred_head [B,N]) in the ContSep flavour, mask alone in the ContExt /
    speechbrain flavours (`returns_pred_head`)."""

    returns_pred_head = True

    def __init__(self, in_channels, out_channels, intra_model, inter_model, num_layers=1, norm="ln", K=200,
                 num_spks=2, skip_around_intra=True, linear_layer_after_inter_intra=True,
                 use_global_pos_enc=False, max_length=20000, llm_dim=4096):
        super().__init__()
        if use_global_pos_enc:
            _unsupported("use_global_pos_enc=True")
        self.K = K
        self.num_spks = num_spks
        self.num_layers = num_layers
        self.norm = select_norm(norm, in_channels, 3)
        self.conv1d = nn.Conv1d(in_channels, out_channels, 1, bias=False)
        self.use_global_pos_enc = use_global_pos_enc
        self.dual_mdl = nn.ModuleList([])
        for _ in range(num_layers):
            self.dual_mdl.append(copy.deepcopy(Dual_Computation_Block_CSE(
                intra_model, inter_model, out_channels, norm, skip_around_intra=skip_around_intra,
                linear_layer_after_inter_intra=linear_layer_after_inter_intra, llm_dim=llm_dim)))
        self.conv2d = nn.Conv2d(out_channels, out_channels * num_spks, kernel_size=1)
        self.end_conv1x1 = nn.Conv1d(out_channels, in_channels, 1, bias=False)
        self.prelu = nn.PReLU()
        self.activation = nn.ReLU()
        self.output = nn.Sequential(nn.Conv1d(out_channels, out_channels, 1), nn.Tanh())
        self.output_gate = nn.Sequential(nn.Conv1d(out_channels, out_channels, 1), nn.Sigmoid())
        self._table = ParamTable()

    def add_ctx(self):
        for i in range(self.num_layers):
            self.dual_mdl[i].add_ctx()

    def _check_supported(self):
        if self.K != CHUNK or self.num_layers != 2 or self.conv1d.weight.shape[0] != N_CH:
            _unsupported(f"Dual_Path_Model(K={self.K}, num_layers={self.num_layers})")
        for blk in self.dual_mdl:
            blk.intra_mdl._check_supported()
            blk.inter_mdl._check_supported()

    def _tensors(self, prefix="masknet."):
        out = {prefix + k: v for k, v in self.named_parameters()}
        out.update({prefix + k: v for k, v in self.named_buffers()})
        return out

    def forward(self, x, ctx=None, precision=None):
        self._check_supported()
        _check_cuda(x, "x")
        _warn_no_grad(self)
        prec = resolve_precision(precision)
        B, N, L = x.shape
        c = 0 if ctx is None else ctx.size(1)
        dev = x.device
        stream = current_stream(dev)
        tensors = self._tensors()
        # the stand-alone masknet never touches encoder/decoder weights; give the table placeholders
        tensors["encoder.conv1d.weight"] = self.conv1d.weight
        tensors["decoder.weight"] = self.conv1d.weight
        params = self._table.build(tensors, self.num_spks, prec, stream)
        E = x.transpose(1, 2).to(_act_dtype(prec)).contiguous()
        lib = _lib.load()
        T_equiv = ENC_S * (L - 1) + ENC_K
        nbytes = lib.cse_workspace_bytes(B, T_equiv, c, self.num_spks, prec)
        if nbytes == 0:
            raise _lib.CseError(lib.cse_last_error().decode())
        keep, ws_ptr, ws_len = WORKSPACE.get(nbytes, dev)
        mask = torch.empty(B, L, self.num_spks, N_CH, dtype=torch.float32, device=dev)
        pred = torch.empty(B, N_CH, dtype=torch.float32, device=dev)
        ctx_d = None if ctx is None else ctx.contiguous().float()
        _lib.call("cse_masknet_fwd", C.byref(params), _lib.ptr(E), _lib.ptr(ctx_d), B, L, c, self.num_spks, prec,
                  _lib.ptr(mask), _lib.ptr(pred), C.c_void_p(ws_ptr), ws_len, C.c_void_p(stream))
        mask = mask.permute(2, 0, 3, 1)                     # [spk,B,N,L] view
        if self.returns_pred_head:
            return mask, pred
        return mask


class Dual_Path_Model(Dual_Path_Model_CSE):
    """speechbrain Dual_Path_Model (sepformer.py:11): no context, forward(x) -> mask."""

    returns_pred_head = False

    def __init__(self, in_channels, out_channels, intra_model, inter_model, num_layers=1, norm="ln", K=200,
                 num_spks=2, skip_around_intra=True, linear_layer_after_inter_intra=True,
                 use_global_pos_enc=False, max_length=20000):
        super().__init__(in_channels, out_channels, intra_model, inter_model, num_layers=num_layers,
                         norm=norm, K=K, num_spks=num_spks, skip_around_intra=skip_around_intra,
                         linear_layer_after_inter_intra=linear_layer_after_inter_intra,
                         use_global_pos_enc=use_global_pos_enc, max_length=max_length, llm_dim=None)

    def forward(self, x, precision=None):
        return super().forward(x, None, precision=precision)


class Dual_Path_Model_CSE_Ext(Dual_Path_Model_CSE):
    """ContExt.py:132-294 flavour: forward returns the mask only (and so do its blocks,
    ContExt.py:557)."""

    returns_pred_head = False

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        for blk in self.dual_mdl:
            blk.returns_pred_head = False


# ----------------------------------------------------------------------------------------------
# fused top-level path shared by the three Sepformer flavours
# ----------------------------------------------------------------------------------------------
def _make_masknet(cls, num_spks, **extra):
    def block():
        return SBTransformerBlock_CSE(num_layers=8, d_model=256, nhead=8, d_ffn=1024, dropout=0,
                                      use_positional_encoding=True, norm_before=True)
    return cls(num_spks=num_spks, in_channels=256, out_channels=256, num_layers=2, K=250,
               intra_model=block(), inter_model=block(), norm="ln",
               linear_layer_after_inter_intra=False, skip_around_intra=True, **extra)


class _SepformerBase(nn.Module):
    """encoder -> masknet(ctx) -> mask * mix_w -> decoder -> pad/trim as ONE C-ABI call."""

    precision = None        # None: follow torch.autocast; or 'fp32' / 'bf16'
    use_cuda_graph = False  # replay each call shape as one CUDA graph (inference)

    max_cached_graphs = 4   # LRU bound on captured call shapes (each holds a full workspace)

    def _init_common(self):
        self._table = ParamTable()
        self._graphs = {}
        self._graph_gen = -1

    def _tensors(self):
        out = {}
        for prefix, mod in (("encoder.", self.encoder), ("masknet.", self.masknet), ("decoder.", self.decoder)):
            for k, v in mod.named_parameters():
                out[prefix + k] = v
            for k, v in mod.named_buffers():
                out[prefix + k] = v
        return out

    def _run(self, mix, ctx, n_masks, want_pred_head):
        """mix [B,T], ctx [B,c,4096] | None -> est [B,T,n_masks] fp32 (+ pred_head [B,256])."""
        self.masknet._check_supported()
        if mix.dim() != 2:
            raise RuntimeError(f"mix must be [B,T], got {tuple(mix.shape)}")
        _check_cuda(mix, "mix")
        prec = resolve_precision(self.precision)
        B, T = mix.shape
        c = 0
        if ctx is not None:
            llm_dim = self.masknet.dual_mdl[0].llm_dim
            if ctx.dim() != 3 or ctx.size(0) != B or ctx.size(2) != llm_dim:
                raise RuntimeError(f"ctx must be [B,c,{llm_dim}], got {tuple(ctx.shape)}")
            _check_cuda(ctx, "ctx")
            c = ctx.size(1)
        if torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters())
                                        or (ctx is not None and ctx.requires_grad)):
            # training step (train_ContSep.py:384-419, train_ContExt.py:366-389): the same stages as autograd
            # nodes whose forward AND backward are C-ABI calls (training.py).  Under torch.autocast (the
            # reference's --fp16 / --bf16 switch, train_ContSep.py:383) the transformer layers run their bf16
            # tensor-core forward / backward; every other stage, the residual stream and the gradients stay fp32
            # — outputs ALWAYS carry the autograd graph (never a silent graph-less result).
            from . import training
            return training.forward_train(self._tensors(), mix, ctx, n_masks, want_pred_head, bf16=(prec == BF16))
        dev = mix.device
        stream = current_stream(dev)
        mix = mix.contiguous().float()
        if ctx is not None:
            ctx = ctx.contiguous().float()
        params = self._table.build(self._tensors(), n_masks, prec, stream)
        lib = _lib.load()
        nbytes = lib.cse_workspace_bytes(B, T, c, n_masks, prec)
        if nbytes == 0:
            raise _lib.CseError(lib.cse_last_error().decode())
        if self.use_cuda_graph:
            return self._run_graph(params, mix, ctx, (B, T, c, n_masks, prec, want_pred_head), nbytes)
        keep, ws_ptr, ws_len = WORKSPACE.get(nbytes, dev)
        est = torch.empty(B, T, n_masks, dtype=torch.float32, device=dev)
        pred = torch.empty(B, N_CH, dtype=torch.float32, device=dev) if want_pred_head else None
        _lib.call("cse_forward", C.byref(params), _lib.ptr(mix), _lib.ptr(ctx), B, T, c, n_masks, prec,
                  _lib.ptr(est), _lib.ptr(pred), C.c_void_p(ws_ptr), ws_len, C.c_void_p(stream))
        return est, pred

    def _run_graph(self, params, mix, ctx, shape_key, nbytes):
        """Replay the ~255 kernel launches of one forward as a single CUDA graph (launch gaps are
        ~6 % of the step at B=16 and dominate at B=1).  One graph per call shape, with its own
        static input / output / workspace buffers; weights are read through the same pointers, and
        the bf16 pack is refreshed outside the graph whenever a parameter version changes."""
        B, T, c, n_masks, prec, want_pred = shape_key
        dev = mix.device
        if self._graph_gen != self._table.generation:      # parameters re-allocated: every baked-in pointer is stale
            self._graphs.clear()
            self._graph_gen = self._table.generation
        key = shape_key + (dev,)
        ent = self._graphs.pop(key, None)
        if ent is not None:
            self._graphs[key] = ent                          # re-insert: most recently used last
        if ent is None:
            while len(self._graphs) >= self.max_cached_graphs:   # each entry holds a workspace + I/O buffers
                self._graphs.pop(next(iter(self._graphs)))
            st = {
                "mix": torch.empty_like(mix),
                "ctx": None if ctx is None else torch.empty_like(ctx),
                "est": torch.empty(B, T, n_masks, dtype=torch.float32, device=dev),
                "pred": torch.empty(B, N_CH, dtype=torch.float32, device=dev) if want_pred else None,
                "ws": torch.empty(nbytes + 256, dtype=torch.uint8, device=dev),
            }
            off = (-st["ws"].data_ptr()) % 256

            def launch():
                _lib.call("cse_forward", C.byref(params), _lib.ptr(st["mix"]), _lib.ptr(st["ctx"]), B, T, c,
                          n_masks, prec, _lib.ptr(st["est"]), _lib.ptr(st["pred"]),
                          C.c_void_p(st["ws"].data_ptr() + off), st["ws"].numel() - off,
                          C.c_void_p(current_stream(dev)))

            st["mix"].copy_(mix)
            if ctx is not None:
                st["ctx"].copy_(ctx)
            launch()                                    # warm-up: one-time attribute / descriptor setup
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                launch()
            ent = (graph, st)
            self._graphs[key] = ent
        graph, st = ent
        st["mix"].copy_(mix)
        if ctx is not None:
            st["ctx"].copy_(ctx)
        graph.replay()
        return st["est"].clone(), (None if st["pred"] is None else st["pred"].clone())

    def separate_host(self, mix_host, ctx_host=None, n_masks=None, est_host=None):
        """End-to-end entry with HOST (pinned) buffers through cse_forward_host: H2D, forward, D2H.
        Returns est_host [B,T,n_masks] (and pred_head_host for the ContSep flavour)."""
        prec = resolve_precision(self.precision)
        dev = next(self.parameters()).device
        stream = current_stream(dev)
        B, T = mix_host.shape
        c = 0 if ctx_host is None else ctx_host.size(1)
        n_masks = n_masks or self._n_masks()
        params = self._table.build(self._tensors(), n_masks, prec, stream)
        lib = _lib.load()
        nbytes = lib.cse_workspace_bytes(B, T, c, n_masks, prec)
        if nbytes == 0:
            raise _lib.CseError(lib.cse_last_error().decode())
        keep, ws_ptr, ws_len = WORKSPACE.get(nbytes, dev)
        if est_host is None:
            est_host = torch.empty(B, T, n_masks, dtype=torch.float32).pin_memory()
        pred_host = torch.empty(B, N_CH, dtype=torch.float32).pin_memory() if self._wants_pred() else None
        _lib.call("cse_forward_host", C.byref(params), _lib.ptr(mix_host), _lib.ptr(ctx_host), B, T, c, n_masks,
                  prec, _lib.ptr(est_host), _lib.ptr(pred_host), C.c_void_p(ws_ptr), ws_len, C.c_void_p(stream))
        return est_host, pred_host

    def host_pipeline(self, B, T, c=0, depth=2, n_masks=None):
        """Pipelined end-to-end entry (cse_pipeline_*): `depth` forwards in flight, H2D of step i+1 and D2H of
        step i-1 overlapping the forward of step i, each forward replayed as one CUDA graph."""
        return HostPipeline(self, B, T, c, n_masks or self._n_masks(), depth)

    def _n_masks(self):
        return self.num_spks

    def _wants_pred(self):
        return False


class HostPipeline:
    """`submit(mix_host, ctx_host)` -> ticket; `wait(ticket)` -> (est_host [B,T,n_masks], pred_head_host | None).
    Host tensors must be pinned; the serving-loop shape of test.py:231-245 with the copies off the critical path."""

    def __init__(self, model, B, T, c, n_masks, depth=2):
        self.model, self.shape, self.depth = model, (B, T, c, n_masks), depth
        prec = resolve_precision(model.precision)
        dev = next(model.parameters()).device
        self._params = model._table.build(model._tensors(), n_masks, prec, current_stream(dev))
        torch.cuda.synchronize(dev)                                   # the bf16 pack is complete before capture
        lib = _lib.load()
        nbytes = lib.cse_pipeline_workspace_bytes(B, T, c, n_masks, prec, depth)
        if nbytes == 0:
            raise _lib.CseError(lib.cse_last_error().decode() or "cse_pipeline_workspace_bytes: bad arguments")
        self._ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
        off = (-self._ws.data_ptr()) % 256
        self._handle = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.call("cse_pipeline_create", C.byref(self._params), B, T, c, n_masks, prec, depth,
                      C.c_void_p(self._ws.data_ptr() + off), self._ws.numel() - off, C.byref(self._handle))
        self._wants_pred = model._wants_pred() and c > 0
        self._est = [torch.empty(B, T, n_masks, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self._pred = [torch.empty(B, N_CH, dtype=torch.float32).pin_memory() if self._wants_pred else None
                      for _ in range(depth)]

    @staticmethod
    def _check_host(t, name, shape):
        if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != shape:
            raise _lib.CseError(f"{name} must be a contiguous float32 HOST tensor of shape {shape}, got "
                                f"{tuple(t.shape)} {t.dtype} on {t.device}")
        if not t.is_pinned():
            raise _lib.CseError(f"{name} must be pinned (page-locked) host memory for the asynchronous copies")

    def submit(self, mix_host, ctx_host=None, est_host=None):
        if self._handle is None:
            raise _lib.CseError("pipeline is closed")
        B, T, c, n_masks = self.shape
        self._check_host(mix_host, "mix_host", (B, T))
        if c:
            self._check_host(ctx_host, "ctx_host", (B, c, _lib.CTX))
        slot = C.c_int(-1)
        # the slot that will be used is the C side's round-robin cursor; its default output buffers are ours
        nxt = getattr(self, "_next", 0)
        est = est_host if est_host is not None else self._est[nxt]
        if est_host is not None:
            self._check_host(est_host, "est_host", (B, T, n_masks))
        pred = self._pred[nxt]
        _lib.call("cse_pipeline_submit", self._handle, _lib.ptr(mix_host), _lib.ptr(ctx_host) if c else None,
                  _lib.ptr(est), _lib.ptr(pred), C.byref(slot))
        self._next = (slot.value + 1) % self.depth
        return (slot.value, est, pred, mix_host, ctx_host)            # keeps the host buffers alive

    def wait(self, ticket):
        _lib.call("cse_pipeline_wait", self._handle, ticket[0])
        return ticket[1], ticket[2]

    def close(self):
        if self._handle is not None:
            _lib.call("cse_pipeline_destroy", self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
