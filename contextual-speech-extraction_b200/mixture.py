"""Loader-side mixture synthesis and collation on the device.

The reference builds every training / evaluation item on CPU workers (`src/data/dataset_train_CSE.py`):
peak-normalise each clip (:237,274), `mix_audio` / `mix_audio_3spk` at a drawn SNR (:257-263,417-505), resample
16 kHz -> 8 kHz (:393-398), then `collate_fn` right-pads to the batch maximum (:507-601).  Here the same steps run on
a whole batch of ragged clips already resident on the GPU; results are the collated `[B, T]` float32 tensors the
training loop consumes (`mixed_sp`, `gt_sp`, `ns_sp_1`, `ns_sp_2`, `sp_len`).  Random draws (which clips, SNRs,
augmentation) stay on the host, as in the reference.  No CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .runtime import current_stream


def _st(dev):
    return C.c_void_p(current_stream(dev))


def _flat(clips, device):
    """list of 1-D tensors / arrays -> (flat float32 CUDA buffer, int64 offsets [B+1] on the device, lengths)."""
    ts = [torch.as_tensor(c, dtype=torch.float32).reshape(-1) for c in clips]
    lens = [int(t.numel()) for t in ts]
    if min(lens) == 0:
        raise RuntimeError("empty clip")
    off = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64)
    flat = torch.cat([t.to(device, non_blocking=True) for t in ts])
    return flat, off.to(device), lens


def _device_of(clips, device):
    if device is not None:
        return torch.device(device)
    for c in clips:
        if isinstance(c, torch.Tensor) and c.is_cuda:
            return c.device
    raise _lib.CseError("mixture synthesis runs on the GPU: pass CUDA tensors or device='cuda' (no CPU fallback)")


def mix_batch(signals, noises, snrs, noises2=None, snrs2=None, pad=True, T_out=None, device=None):
    """Batched `mix_audio` (noises2 is None) or `mix_audio_3spk`, fused with `collate_fn`'s right padding.

    signals / noises / noises2: lists of B 1-D clips (ragged); snrs / snrs2: B values (the dataset draws
    `np.clip(random.normalvariate(0, 4), -5, 5)`).  Returns (mixed, signal, noise1[, noise2]) as [B, T_out] float32
    CUDA tensors and sp_len [B] int32 — the tuple order of the reference's functions."""
    dev = _device_of(list(signals) + list(noises), device)
    if dev.type != "cuda":
        raise _lib.CseError(f"mixture synthesis on {dev}: the CUDA path has no CPU fallback")
    B = len(signals)
    if len(noises) != B or len(snrs) != B or (noises2 is not None and (len(noises2) != B or len(snrs2) != B)):
        raise RuntimeError("mix_batch: every list needs one entry per item")
    s_flat, s_off, s_len = _flat(signals, dev)
    a_flat, a_off, a_len = _flat(noises, dev)
    three = noises2 is not None
    if three:
        c_flat, c_off, c_len = _flat(noises2, dev)
        lens = [max(x) for x in zip(s_len, a_len, c_len)]
    else:
        lens = s_len
    T_out = int(T_out or max(lens))
    if T_out < max(lens):
        raise RuntimeError(f"T_out={T_out} is shorter than the longest mixture ({max(lens)})")
    snr1 = torch.tensor([float(s) for s in snrs], dtype=torch.float64).to(dev)
    snr2 = torch.tensor([float(s) for s in snrs2], dtype=torch.float64).to(dev) if three else None
    outs = [torch.empty(B, T_out, dtype=torch.float32, device=dev) for _ in range(4 if three else 3)]
    sp_len = torch.empty(B, dtype=torch.int32, device=dev)
    _lib.call("cse_mix_audio", _lib.ptr(s_flat), _lib.ptr(s_off), _lib.ptr(a_flat), _lib.ptr(a_off),
              _lib.ptr(c_flat) if three else None, _lib.ptr(c_off) if three else None, _lib.ptr(snr1), _lib.ptr(snr2),
              B, 2 if three else 1, int(bool(pad)), T_out, _lib.ptr(outs[0]), _lib.ptr(outs[1]), _lib.ptr(outs[2]),
              _lib.ptr(outs[3]) if three else None, _lib.ptr(sp_len), _st(dev))
    return tuple(outs) + (sp_len,)


def peak_normalize(clips, peak=0.9, T_out=None, device=None):
    """`x / np.max(np.abs(x)) * 0.9` (dataset_train_CSE.py:237,274) for a batch of ragged clips -> ([B, T_out], lengths)."""
    dev = _device_of(list(clips), device)
    flat, off, lens = _flat(clips, dev)
    T_out = int(T_out or max(lens))
    if T_out < max(lens):
        raise RuntimeError(f"T_out={T_out} is shorter than the longest clip ({max(lens)})")
    out = torch.empty(len(lens), T_out, dtype=torch.float32, device=dev)
    _lib.call("cse_peak_normalize", _lib.ptr(flat), _lib.ptr(off), len(lens), float(peak), T_out, _lib.ptr(out), _st(dev))
    return out, lens


def kaiser_lowpass_taps(down, beta=5.0):
    """The FIR scipy.signal.resample_poly(x, 1, down) designs: 20 * down + 1 taps, cut-off at the new Nyquist,
    Kaiser(beta = 5) window, unit DC gain."""
    half = 10 * down
    n = np.arange(2 * half + 1) - half
    fc = 1.0 / down
    h = fc * np.sinc(fc * n) * np.kaiser(2 * half + 1, beta)
    return (h / h.sum()).astype(np.float32)


def decimate(x, lengths=None, down=2, taps=None):
    """The 16 kHz -> 8 kHz step (dataset_train_CSE.py:393-398) for a collated batch: x [B, T_in] CUDA float32,
    lengths [B] int32 (valid samples per row) or None -> (y [B, ceil(T_in / down)], new lengths [B] int32).
    `taps`: any odd-length low-pass (centre tap aligned with the kept samples); default `kaiser_lowpass_taps(down)`."""
    if not x.is_cuda:
        raise _lib.CseError(f"x is on {x.device}: the CUDA path has no CPU fallback")
    x = x.float().contiguous()
    B, T_in = x.shape
    h = torch.as_tensor(kaiser_lowpass_taps(down) if taps is None else np.asarray(taps, dtype=np.float32)).to(x.device)
    if h.numel() % 2 != 1:
        raise RuntimeError("decimate: the filter needs an odd number of taps")
    T_out = -(-T_in // down)
    y = torch.empty(B, T_out, dtype=torch.float32, device=x.device)
    new_len = torch.empty(B, dtype=torch.int32, device=x.device)
    lens = None if lengths is None else lengths.to(device=x.device, dtype=torch.int32).contiguous()
    _lib.call("cse_decimate", _lib.ptr(x), _lib.ptr(lens), B, T_in, int(down), _lib.ptr(h), h.numel(), T_out,
              _lib.ptr(y), _lib.ptr(new_len), _st(x.device))
    return y, new_len
