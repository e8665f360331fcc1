"""Deterministic synthetic weights and inputs.

There are no checkpoints or corpora (no network), so every parity test and benchmark runs on
random-init weights of the reference architecture and synthetic 8 kHz mixtures.  Values come
from numpy's PCG64 streams keyed by (seed, crc32(name)), so they are identical on every box
and independent of torch's RNG.  EVERY parameter is drawn independently (incl. norm affines,
biases, PReLU): the reference deep-copies one dual block (ContSep.py:171-185), which would
otherwise hide block-index bugs.
"""
import math
import zlib

import numpy as np
import torch

from .shapes import (CTX_DIM, D_FFN, ENC_K, N_BLOCK, N_CH, N_LAYER, PE_MAX, SE_DIM)

VARIANTS = ("sepformer", "contsep", "context", "hcontext")


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))


def positional_table(max_len: int = PE_MAX, d: int = N_CH) -> torch.Tensor:
    """Sinusoid table [1, max_len, d] (speechbrain PositionalEncoding; CSE_transformer.py:88)."""
    pe = torch.zeros(max_len, d)
    pos = torch.arange(0, max_len).unsqueeze(1).float()
    freq = torch.exp(torch.arange(0, d, 2).float() * -(math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(pos * freq)
    pe[:, 1::2] = torch.cos(pos * freq)
    return pe.unsqueeze(0)


def param_spec(variant: str = "contsep", num_spks: int = 2, ce: bool = True):
    """Ordered {state_dict key: (shape, kind)} for a reference model variant.

    Key names/shapes are the compatibility contract: `load_state_dict(strict=True)`
    into the reference's own modules must succeed with these.
    kind: 'w' fan-in-scaled weight, 'b' small bias, 'g' norm gain ~1, 'prelu', 'pe' buffer.
    """
    if variant not in VARIANTS:
        raise ValueError(f"unknown variant {variant!r}; expected one of {VARIANTS}")
    N, F = N_CH, D_FFN
    spec = {}
    spec["encoder.conv1d.weight"] = ((N, 1, ENC_K), "w")
    spec["masknet.norm.weight"] = ((N,), "g")
    spec["masknet.norm.bias"] = ((N,), "b")
    spec["masknet.conv1d.weight"] = ((N, N, 1), "w")
    with_ctx = variant != "sepformer"
    for i in range(N_BLOCK):
        for path in ("intra", "inter"):
            p = f"masknet.dual_mdl.{i}.{path}_mdl."
            for l in range(N_LAYER):
                q = f"{p}mdl.layers.{l}."
                spec[q + "self_att.att.in_proj_weight"] = ((3 * N, N), "w")
                spec[q + "self_att.att.in_proj_bias"] = ((3 * N,), "b")
                spec[q + "self_att.att.out_proj.weight"] = ((N, N), "w")
                spec[q + "self_att.att.out_proj.bias"] = ((N,), "b")
                spec[q + "pos_ffn.ffn.0.weight"] = ((F, N), "w")
                spec[q + "pos_ffn.ffn.0.bias"] = ((F,), "b")
                spec[q + "pos_ffn.ffn.3.weight"] = ((N, F), "w")
                spec[q + "pos_ffn.ffn.3.bias"] = ((N,), "b")
                spec[q + "norm1.norm.weight"] = ((N,), "g")
                spec[q + "norm1.norm.bias"] = ((N,), "b")
                spec[q + "norm2.norm.weight"] = ((N,), "g")
                spec[q + "norm2.norm.bias"] = ((N,), "b")
            spec[p + "mdl.norm.norm.weight"] = ((N,), "g")
            spec[p + "mdl.norm.norm.bias"] = ((N,), "b")
            spec[p + "pos_enc.pe"] = ((1, PE_MAX, N), "pe")
        d = f"masknet.dual_mdl.{i}."
        spec[d + "intra_norm.weight"] = ((N,), "g")
        spec[d + "intra_norm.bias"] = ((N,), "b")
        spec[d + "inter_norm.weight"] = ((N,), "g")
        spec[d + "inter_norm.bias"] = ((N,), "b")
        if with_ctx:
            spec[d + "intra_context_mapper.weight"] = ((N, CTX_DIM), "w")
            spec[d + "intra_context_mapper.bias"] = ((N,), "b")
            spec[d + "inter_context_mapper.weight"] = ((N, CTX_DIM), "w")
            spec[d + "inter_context_mapper.bias"] = ((N,), "b")
    spec["masknet.conv2d.weight"] = ((N * num_spks, N, 1, 1), "w")
    spec["masknet.conv2d.bias"] = ((N * num_spks,), "b")
    spec["masknet.end_conv1x1.weight"] = ((N, N, 1), "w")
    spec["masknet.prelu.weight"] = ((1,), "prelu")
    spec["masknet.output.0.weight"] = ((N, N, 1), "w")
    spec["masknet.output.0.bias"] = ((N,), "b")
    spec["masknet.output_gate.0.weight"] = ((N, N, 1), "w")
    spec["masknet.output_gate.0.bias"] = ((N,), "b")
    spec["decoder.weight"] = ((N, 1, ENC_K), "w")
    if variant == "contsep":
        n_sel = 1 if (num_spks == 2 and not ce) else num_spks      # ContSep.py:48-51
        spec["context_selector.weight"] = ((n_sel, N), "w")
        spec["context_selector.bias"] = ((n_sel,), "b")
    if variant == "hcontext":
        spec["se_embedding.weight"] = ((CTX_DIM, SE_DIM), "w")       # ContExt.py:51-52
        spec["se_embedding.bias"] = ((CTX_DIM,), "b")
    return spec


def make_state_dict(variant: str = "contsep", num_spks: int = 2, seed: int = 0, ce: bool = True,
                    dtype=torch.float32):
    """Independent random value for every parameter of `variant` (CPU tensors)."""
    out = {}
    for name, (shape, kind) in param_spec(variant, num_spks, ce).items():
        if kind == "pe":
            out[name] = positional_table().to(dtype)
            continue
        g = _rng(seed, name)
        if kind == "w":
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            if name == "decoder.weight":            # ConvTranspose1d: in_channels is dim 0
                fan_in = shape[0]
            v = g.standard_normal(shape) / math.sqrt(fan_in)
        elif kind == "b":
            v = 0.1 * g.standard_normal(shape)
        elif kind == "g":
            v = 1.0 + 0.1 * g.standard_normal(shape)
        elif kind == "prelu":
            v = 0.25 + 0.05 * g.standard_normal(shape)
        else:  # pragma: no cover
            raise AssertionError(kind)
        out[name] = torch.from_numpy(np.ascontiguousarray(v)).to(dtype)
    return out


def _lowpass_noise(g: np.random.Generator, n: int, pole: float) -> np.ndarray:
    x = g.standard_normal(n)
    a = 1.0 - pole
    # 1-pole low-pass; vectorised through the closed form of the IIR impulse response
    k = int(min(n, max(16, math.ceil(math.log(1e-6) / math.log(max(pole, 1e-6))))))
    h = a * pole ** np.arange(k)
    return np.convolve(x, h)[:n]


def make_sources(B: int, T: int, n_src: int = 2, seed: int = 1234) -> torch.Tensor:
    """[B, T, n_src] band-limited noise sources, each peak-normalised to 0.9
    (dataset_train_CSE.py:237)."""
    out = np.zeros((B, T, n_src), dtype=np.float64)
    for b in range(B):
        for s in range(n_src):
            g = _rng(seed, f"src.{b}.{s}")
            y = _lowpass_noise(g, T, pole=0.6 + 0.3 * g.random())
            out[b, :, s] = 0.9 * y / max(np.abs(y).max(), 1e-9)
    return torch.from_numpy(out).float()


def make_mixture(B: int, T: int, n_src: int = 2, seed: int = 1234):
    """Mix the sources at snr = clip(N(0,4), -5, 5) dB relative to source 0
    (dataset_train_CSE.py:257,417-456) and rescale to peak 0.9.
    Returns (mix [B,T], sources [B,T,n_src] scaled consistently)."""
    src = make_sources(B, T, n_src, seed).double().numpy()
    mix = np.zeros((B, T))
    for b in range(B):
        g = _rng(seed, f"snr.{b}")
        ref_pow = np.mean(src[b, :, 0] ** 2) + 1e-12
        for s in range(1, n_src):
            snr = float(np.clip(g.normal(0.0, 4.0), -5.0, 5.0))
            p = np.mean(src[b, :, s] ** 2) + 1e-12
            src[b, :, s] *= math.sqrt(ref_pow / (p * 10 ** (snr / 10)))
        m = src[b].sum(-1)
        scale = 0.9 / max(np.abs(m).max(), 1e-9)
        mix[b] = m * scale
        src[b] *= scale
    return torch.from_numpy(mix).float(), torch.from_numpy(src).float()


def make_context(B: int, c: int = 1, seed: int = 1234, dim: int = CTX_DIM) -> torch.Tensor:
    """[B, c, dim] stand-in for Llama-3-8B last-token hidden states (RMS ~ 1 per element)."""
    g = _rng(seed, f"ctx.{c}.{dim}")
    return torch.from_numpy(g.standard_normal((B, c, dim))).float()


def make_speaker_embedding(B: int, seed: int = 1234) -> torch.Tensor:
    """[B, 1, 192] stand-in for the ECAPA voice cue (train_HContExt.py:367)."""
    g = _rng(seed, "se")
    return torch.from_numpy(g.standard_normal((B, 1, SE_DIM))).float()
