"""TEST INFRASTRUCTURE ONLY — CPU restatement of the loader-side mixture synthesis.

  mix_audio / mix_audio_3spk   src/data/dataset_train_CSE.py:417-505 (identical to mix_aud.py:3-96), restated
                               line by line in numpy with the same dtype promotion (float32 energies, float64
                               gains / mixture / scale).  PINNED: tests/golden/mixture_*.npz are outputs of the
                               reference's own mix_aud.py (pure numpy, importable here) on seeded inputs
                               (tests/golden/make_golden_mixture.py).
  peak_normalize               dataset_train_CSE.py:237,274  x / max|x| * 0.9 in float32
  collate                      dataset_train_CSE.py:507-601  right padding to the batch maximum
  decimate                     dataset_train_CSE.py:393-398 calls librosa.resample(16 k -> 8 k) whose default
                               back end (soxr_hq) is an absent third-party library: PARITY UNPINNED for the
                               resampler.  The kernel takes its FIR taps as an argument; the default taps and
                               this restatement follow scipy.signal.resample_poly (up = 1; Kaiser beta = 5 design,
                               scipy/signal/_signaltools.py), pinned against scipy itself in tests/test_mixture.py.
Never imported by the product path."""
import numpy as np


def mix_audio(signal, noise, snr, pad=False):
    if not pad and len(signal) > len(noise):
        noise = noise[np.arange(len(signal)) % len(noise)]
    if len(signal) < len(noise):
        noise = noise[:len(signal)]
    noise = noise.astype(np.float32)
    signal = signal.astype(np.float32)
    signal_energy = np.mean(signal ** 2)
    noise_energy = np.mean(noise ** 2)
    g = np.sqrt(10.0 ** (-snr / 10) * signal_energy / noise_energy)
    a = np.sqrt(1 / (1 + g ** 2))
    b = np.sqrt(g ** 2 / (1 + g ** 2))
    if pad and len(signal) > len(noise):
        noise = np.concatenate([noise, np.zeros(len(signal) - len(noise))], 0)
    signal = a * signal
    noise = b * noise
    mixed = signal + noise
    scale = 1 / np.max(np.abs(mixed)) * 0.9
    return scale * mixed, scale * signal, scale * noise


def mix_audio_3spk(signal, noise1, noise2, snr1, snr2, pad=False):
    max_len = max(len(signal), len(noise1), len(noise2))
    if not pad:
        if max_len > len(signal):
            signal = signal[np.arange(max_len) % len(signal)]
        if max_len > len(noise1):
            noise1 = noise1[np.arange(max_len) % len(noise1)]
        if max_len > len(noise2):
            noise2 = noise2[np.arange(max_len) % len(noise2)]
    noise1, noise2, signal = noise1.astype(np.float32), noise2.astype(np.float32), signal.astype(np.float32)
    e_s, e_1, e_2 = np.mean(signal ** 2), np.mean(noise1 ** 2), np.mean(noise2 ** 2)
    g1 = np.sqrt(10.0 ** (-snr1 / 10) * e_s / e_1)
    g2 = np.sqrt(10.0 ** (-snr2 / 10) * e_s / e_2)
    if pad:
        z = lambda x: np.concatenate([x, np.zeros(max_len - len(x))], 0) if max_len > len(x) else x  # noqa: E731
        signal, noise1, noise2 = z(signal), z(noise1), z(noise2)
    noise1 = g1 * noise1
    noise2 = g2 * noise2
    mixed = signal + noise1 + noise2
    scale = 1 / np.max(np.abs(mixed)) * 0.9
    return scale * mixed, scale * signal, scale * noise1, scale * noise2


def peak_normalize(x, peak=0.9):
    x = np.asarray(x, dtype=np.float32)
    return x / np.max(np.abs(x)) * peak            # float32 throughout (python float is weak under NEP 50)


def collate(rows):
    """Right-pad to the batch maximum and stack as float32 (dataset_train_CSE.py:507-601)."""
    n = max(len(r) for r in rows)
    return np.array([np.concatenate([r, np.zeros(n - len(r))], 0) for r in rows]).astype(np.float32)


def kaiser_lowpass_taps(down, beta=5.0):
    """scipy.signal.resample_poly's default filter for up = 1: firwin(2 * 10 * down + 1, 1 / down, ('kaiser', beta))."""
    half = 10 * down
    n = np.arange(2 * half + 1) - half
    fc = 1.0 / down
    h = fc * np.sinc(fc * n) * np.kaiser(2 * half + 1, beta)
    return h / h.sum()


def decimate(x, down, taps=None):
    """y[i] = sum_j h[j] x[i * down + len(h) // 2 - j], zeros outside: scipy.signal.resample_poly(x, 1, down)."""
    x = np.asarray(x, dtype=np.float64)
    h = kaiser_lowpass_taps(down) if taps is None else np.asarray(taps, dtype=np.float64)
    full = np.convolve(x, h)                       # full[k] = sum_j h[j] x[k - j]
    n_out = -(-len(x) // down)
    idx = np.arange(n_out) * down + len(h) // 2
    return full[idx]
