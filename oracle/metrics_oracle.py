"""TEST INFRASTRUCTURE ONLY — CPU restatement of the evaluation metrics of the reference's test loop.

The reference (test.py:198-201,241-245,291-301) uses torchmetrics.audio.{ScaleInvariantSignalNoiseRatio,
SignalDistortionRatio}.  torchmetrics is an un-vendored, un-pinned dependency (README.md:21) that is NOT installable
here, so — like the speechbrain leaves (oracle/sb_shim) — its published algorithm (torchmetrics 1.x,
functional/audio/sdr.py: `signal_distortion_ratio`, `_compute_autocorr_crosscorr`, `_symmetric_toeplitz`) is restated
below in numpy float64: PARITY UNPINNED against torchmetrics itself; pinned against an independent dense
time-domain evaluation of the same definition (tests/test_metrics.py).  Never imported by the product path."""
import math

import numpy as np


def signal_distortion_ratio(preds, target, filter_length=512, zero_mean=False, load_diag=None):
    """preds, target [..., T] -> SDR in dB (float64), torchmetrics' FFT formulation + dense solve."""
    preds = np.asarray(preds, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    if zero_mean:
        preds = preds - preds.mean(-1, keepdims=True)
        target = target - target.mean(-1, keepdims=True)
    target = target / np.maximum(np.linalg.norm(target, axis=-1, keepdims=True), 1e-6)
    preds = preds / np.maximum(np.linalg.norm(preds, axis=-1, keepdims=True), 1e-6)
    n_fft = 2 ** math.ceil(math.log2(preds.shape[-1] + target.shape[-1] - 1))
    t_fft = np.fft.rfft(target, n=n_fft, axis=-1)
    r_0 = np.fft.irfft(t_fft.real ** 2 + t_fft.imag ** 2, n=n_fft)[..., :filter_length]
    p_fft = np.fft.rfft(preds, n=n_fft, axis=-1)
    b = np.fft.irfft(np.conj(t_fft) * p_fft, n=n_fft, axis=-1)[..., :filter_length]
    if load_diag is not None:
        r_0 = r_0.copy()
        r_0[..., 0] += load_diag
    idx = np.abs(np.arange(filter_length)[:, None] - np.arange(filter_length)[None, :])
    r = r_0[..., idx]                                   # symmetric Toeplitz
    sol = np.linalg.solve(r, b[..., None])[..., 0]
    coh = np.einsum("...l,...l->...", b, sol)
    return 10.0 * np.log10(coh / (1.0 - coh))


def signal_distortion_ratio_time_domain(preds, target, filter_length=512):
    """Independent check of the definition: SDR = energy of the best length-L FIR projection of preds onto shifted
    copies of target vs the residual, solved by least squares on the explicit (T+L-1) x L convolution matrix."""
    p = np.asarray(preds, dtype=np.float64)
    t = np.asarray(target, dtype=np.float64)
    t = t / max(np.linalg.norm(t), 1e-6)
    p = p / max(np.linalg.norm(p), 1e-6)
    T, L = len(t), filter_length
    # columns = target delayed by k samples: the cross-correlation b[k] = sum_t target[t] preds[t+k] is A[:, k] . preds
    A = np.zeros((T, L))
    for k in range(L):
        A[k:, k] = t[:T - k]
    G = A.T @ A          # != Toeplitz autocorrelation at the edges: torchmetrics uses the infinite-support form
    r0 = np.array([t[:T - k] @ t[k:] for k in range(L)])
    idx = np.abs(np.arange(L)[:, None] - np.arange(L)[None, :])
    R = r0[idx]
    b = A.T @ p
    sol = np.linalg.solve(R, b)
    coh = b @ sol
    return 10.0 * np.log10(coh / (1.0 - coh)), np.abs(G - R).max()


def scale_invariant_signal_noise_ratio(preds, target):
    """torchmetrics functional SI-SNR (zero_mean=True inside the metric), float64."""
    p = np.asarray(preds, dtype=np.float64)
    t = np.asarray(target, dtype=np.float64)
    eps = np.finfo(np.float32).eps
    p = p - p.mean(-1, keepdims=True)
    t = t - t.mean(-1, keepdims=True)
    alpha = ((p * t).sum(-1, keepdims=True) + eps) / ((t ** 2).sum(-1, keepdims=True) + eps)
    ts = alpha * t
    noise = ts - p
    return 10 * np.log10(((ts ** 2).sum(-1) + eps) / ((noise ** 2).sum(-1) + eps))


class RunningMean:
    """sum / total state of a torchmetrics metric object."""

    def __init__(self):
        self.s, self.n = 0.0, 0

    def update(self, values):
        v = np.asarray(values, dtype=np.float64).reshape(-1)
        self.s += float(v.sum())
        self.n += v.size

    def compute(self):
        return self.s / self.n
