"""Host-side runtime shared by the module mirror: parameter marshalling, bf16 weight pack cache,
workspace cache, precision policy.  PyTorch supplies device memory and streams only."""
import ctypes as C

import torch

from . import _lib
from ._lib import BF16, FP32, BLOCKS, LAYERS


def _dev_f32(t, name):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.CseError(f"{name} is on {t.device}; the CUDA path has no CPU fallback (call .cuda())")
    if t.dtype != torch.float32:
        raise _lib.CseError(f"{name} has dtype {t.dtype}; parameters must be float32 masters")
    if not t.is_contiguous():
        raise _lib.CseError(f"{name} is not contiguous")
    return C.c_void_p(t.data_ptr())


def resolve_precision(explicit=None):
    """fp32 unless the caller runs under torch.autocast (the reference's --fp16/--bf16 switch,
    train_ContSep.py:383) or forces it.  Both autocast dtypes map to the bf16 tensor-core mode."""
    if explicit in ("fp32", FP32):
        return FP32
    if explicit in ("bf16", BF16):
        return BF16
    if explicit is not None:
        raise ValueError(f"unknown precision {explicit!r} (expected 'fp32' or 'bf16')")
    return BF16 if torch.is_autocast_enabled() else FP32


class GraphedStep:
    """A whole training step — `fn(*static_args)`: forward, loss, backward, clip + optimiser — captured into ONE CUDA
    graph and replayed.  The autocast step of BASELINE configs[2] is ~1000 launches of 5-25 us each, so run eagerly it
    is bound by the launching host thread (18-33 ms on a shared box); replayed it runs at the GPU's pace.

        opt = AdamW(model.parameters(), lr=1e-4, amsgrad=True)
        def step(mix, ctx, tgt):
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = -si_snr(model(mix, ctx)[:, :, 0], tgt)
            loss.backward()
            opt.step(max_norm=5.0)
            return loss
        graphed = GraphedStep(step, mix_static, ctx_static, tgt_static)
        loss = graphed(mix, ctx, tgt)         # copies the arguments into the static tensors, replays

    Requirements (the usual ones of torch.cuda.graph): fixed shapes, no host synchronisation inside `fn`, the fused
    optimiser of this package (or a capturable torch optimiser); `fn` runs `warmup` times eagerly first.  Single
    process per graph: a DistributedDataParallel step is not captured here."""

    def __init__(self, fn, *static_args, warmup=3, stream=None):
        self.args = static_args
        side = stream if stream is not None else torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(*static_args)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn(*static_args)

    def __call__(self, *args):
        if len(args) != len(self.args):
            raise ValueError(f"GraphedStep: expected {len(self.args)} arguments, got {len(args)}")
        for dst, src in zip(self.args, args):
            if src is not dst:
                if src.shape != dst.shape or src.dtype != dst.dtype:
                    raise _lib.CseError(f"GraphedStep: argument {tuple(src.shape)} {src.dtype} does not match the "
                                        f"captured {tuple(dst.shape)} {dst.dtype}")
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.out


class ParamTable:
    """Builds the `cse_params` struct from a module's state and keeps the bf16 pack fresh."""

    def __init__(self):
        self._key = None
        self._params = None
        self._pack = None
        self._pack_key = None
        self._keep = None
        self.generation = 0      # bumped whenever the pointer table is rebuilt (CUDA-graph caches key on it)

    @staticmethod
    def _stack(sp, prefix, sd):
        for l in range(LAYERS):
            q = f"{prefix}mdl.layers.{l}."
            lp = sp.layer[l]
            lp.in_proj_w = _dev_f32(sd[q + "self_att.att.in_proj_weight"], q + "in_proj_weight")
            lp.in_proj_b = _dev_f32(sd[q + "self_att.att.in_proj_bias"], q + "in_proj_bias")
            lp.out_proj_w = _dev_f32(sd[q + "self_att.att.out_proj.weight"], q + "out_proj.weight")
            lp.out_proj_b = _dev_f32(sd[q + "self_att.att.out_proj.bias"], q + "out_proj.bias")
            lp.ffn1_w = _dev_f32(sd[q + "pos_ffn.ffn.0.weight"], q + "ffn.0.weight")
            lp.ffn1_b = _dev_f32(sd[q + "pos_ffn.ffn.0.bias"], q + "ffn.0.bias")
            lp.ffn2_w = _dev_f32(sd[q + "pos_ffn.ffn.3.weight"], q + "ffn.3.weight")
            lp.ffn2_b = _dev_f32(sd[q + "pos_ffn.ffn.3.bias"], q + "ffn.3.bias")
            lp.ln1_g = _dev_f32(sd[q + "norm1.norm.weight"], q + "norm1")
            lp.ln1_b = _dev_f32(sd[q + "norm1.norm.bias"], q + "norm1")
            lp.ln2_g = _dev_f32(sd[q + "norm2.norm.weight"], q + "norm2")
            lp.ln2_b = _dev_f32(sd[q + "norm2.norm.bias"], q + "norm2")
        sp.final_g = _dev_f32(sd[prefix + "mdl.norm.norm.weight"], prefix + "mdl.norm")
        sp.final_b = _dev_f32(sd[prefix + "mdl.norm.norm.bias"], prefix + "mdl.norm")
        sp.pe = _dev_f32(sd[prefix + "pos_enc.pe"], prefix + "pos_enc.pe")

    def build(self, tensors, n_masks, precision, stream):
        """tensors: {reference state_dict key -> tensor} for encoder./masknet./decoder. entries.
        Returns a ctypes Params whose *_bf16 members are valid when precision == BF16."""
        key = tuple((k, t.data_ptr()) for k, t in tensors.items())
        if key != self._key:
            p = _lib.Params()
            sd = tensors
            p.enc_w = _dev_f32(sd["encoder.conv1d.weight"], "encoder.conv1d.weight")
            p.norm_g = _dev_f32(sd["masknet.norm.weight"], "masknet.norm.weight")
            p.norm_b = _dev_f32(sd["masknet.norm.bias"], "masknet.norm.bias")
            p.conv1d_w = _dev_f32(sd["masknet.conv1d.weight"], "masknet.conv1d.weight")
            for i in range(BLOCKS):
                d = f"masknet.dual_mdl.{i}."
                bp = p.block[i]
                self._stack(bp.intra, d + "intra_mdl.", sd)
                self._stack(bp.inter, d + "inter_mdl.", sd)
                bp.intra_norm_g = _dev_f32(sd[d + "intra_norm.weight"], d + "intra_norm")
                bp.intra_norm_b = _dev_f32(sd[d + "intra_norm.bias"], d + "intra_norm")
                bp.inter_norm_g = _dev_f32(sd[d + "inter_norm.weight"], d + "inter_norm")
                bp.inter_norm_b = _dev_f32(sd[d + "inter_norm.bias"], d + "inter_norm")
                bp.intra_map_w = _dev_f32(sd.get(d + "intra_context_mapper.weight"), d + "intra_context_mapper")
                bp.intra_map_b = _dev_f32(sd.get(d + "intra_context_mapper.bias"), d + "intra_context_mapper")
                bp.inter_map_w = _dev_f32(sd.get(d + "inter_context_mapper.weight"), d + "inter_context_mapper")
                bp.inter_map_b = _dev_f32(sd.get(d + "inter_context_mapper.bias"), d + "inter_context_mapper")
            p.prelu = _dev_f32(sd["masknet.prelu.weight"], "masknet.prelu.weight")
            p.conv2d_w = _dev_f32(sd["masknet.conv2d.weight"], "masknet.conv2d.weight")
            p.conv2d_b = _dev_f32(sd["masknet.conv2d.bias"], "masknet.conv2d.bias")
            p.out_w = _dev_f32(sd["masknet.output.0.weight"], "masknet.output.0.weight")
            p.out_b = _dev_f32(sd["masknet.output.0.bias"], "masknet.output.0.bias")
            p.gate_w = _dev_f32(sd["masknet.output_gate.0.weight"], "masknet.output_gate.0.weight")
            p.gate_b = _dev_f32(sd["masknet.output_gate.0.bias"], "masknet.output_gate.0.bias")
            p.end_w = _dev_f32(sd["masknet.end_conv1x1.weight"], "masknet.end_conv1x1.weight")
            p.dec_w = _dev_f32(sd["decoder.weight"], "decoder.weight")
            for d in (f"masknet.dual_mdl.{i}." for i in range(BLOCKS)):
                for k in ("intra_context_mapper.weight", "inter_context_mapper.weight"):
                    w = sd.get(d + k)
                    if w is not None and tuple(w.shape) != (_lib.N, _lib.CTX):
                        raise _lib.CseError(f"{d + k} has shape {tuple(w.shape)}, the kernels need ({_lib.N}, {_lib.CTX})")
            self._params, self._key = p, key
            self.generation += 1
            self._pack_key = None
            self._keep = list(tensors.values())     # keep storages alive while pointers are cached
        if precision == BF16:
            pack_key = (n_masks,) + tuple(t._version for t in tensors.values())
            if pack_key != self._pack_key:
                n = _lib.load().cse_pack_bf16_elems(n_masks)
                dev = tensors["encoder.conv1d.weight"].device
                if self._pack is None or self._pack.numel() < n or self._pack.device != dev:
                    self._pack = torch.empty(n, dtype=torch.bfloat16, device=dev)
                _lib.call("cse_pack_bf16", C.byref(self._params), n_masks, _lib.ptr(self._pack), n,
                          C.c_void_p(stream))
                self._pack_key = pack_key
        return self._params


class Workspace:
    """Grow-only scratch buffer per device (caller-owned memory of the C ABI)."""

    def __init__(self):
        self._buf = {}

    def get(self, nbytes, device):
        buf = self._buf.get(device)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
            self._buf[device] = buf
        off = (-buf.data_ptr()) % 256
        return buf, buf.data_ptr() + off, buf.numel() - off


WORKSPACE = Workspace()


def current_stream(device):
    return torch.cuda.current_stream(device).cuda_stream
