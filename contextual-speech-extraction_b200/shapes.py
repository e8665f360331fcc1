"""Shape algebra of the Sepformer dual-path hot path (host logic, no torch needed).

Reference: Encoder = Conv1d(k=16, s=8) (ContSep.py:10), chunking
`Dual_Path_Model_CSE._padding/_Segmentation` (ContSep.py:270-335), ConvTranspose1d decoder
(ContSep.py:40).  All constants are hard-coded in the reference constructors
(ContSep.py:10-40); the kernels are specialised to them.
"""
from dataclasses import dataclass

N_CH = 256        # encoder filters / model width          (ContSep.py:10,13-14,19)
ENC_K = 16        # encoder kernel                         (ContSep.py:10)
ENC_S = 8         # encoder stride = kernel // 2           (speechbrain Encoder)
CHUNK = 250       # K, chunk length                        (ContSep.py:16)
HOP = CHUNK // 2  # P, 50 % overlap                        (ContSep.py:286)
N_HEAD = 8        # heads                                  (ContSep.py:20)
D_HEAD = N_CH // N_HEAD
D_FFN = 1024      # FFN hidden                             (ContSep.py:21)
N_LAYER = 8       # transformer layers per stack           (ContSep.py:18)
N_BLOCK = 2       # dual-path blocks                       (ContSep.py:15)
CTX_DIM = 4096    # Llama-3-8B hidden                      (ContSep.py:8)
SE_DIM = 192      # ECAPA embedding                        (ContExt.py:52)
PE_MAX = 2500     # speechbrain PositionalEncoding max_len (CSE_transformer.py:88)
SAMPLE_RATE = 8000


@dataclass(frozen=True)
class PathShape:
    """Every derived length for one call of the hot path."""
    B: int        # mixtures
    T: int        # samples per mixture
    c: int        # context tokens (0 plain Sepformer, 1 ContSep/ContExt, 2 HContExt)
    spk: int      # masks estimated
    L: int        # encoder frames
    gap: int      # right zero padding added by _padding
    S: int        # chunks
    T_est: int    # decoder output length before the pad/trim fix (ContSep.py:90-95)

    @property
    def n_intra(self):  # tokens per intra sequence
        return CHUNK + self.c

    @property
    def n_inter(self):  # tokens per inter sequence
        return self.S + self.c

    @property
    def rows_intra(self):
        return self.B * self.S * self.n_intra

    @property
    def rows_inter(self):
        return self.B * CHUNK * self.n_inter

    @property
    def audio_seconds(self):
        return self.B * self.T / SAMPLE_RATE


def path_shape(B: int, T: int, c: int = 1, spk: int = 2) -> PathShape:
    if T < ENC_K:
        raise ValueError(f"mixture of {T} samples is shorter than the {ENC_K}-tap encoder kernel")
    if B < 1 or c < 0 or spk < 1:
        raise ValueError("B >= 1, c >= 0, spk >= 1 required")
    L = (T - ENC_K) // ENC_S + 1
    gap = CHUNK - (HOP + L % CHUNK) % CHUNK          # ContSep.py:287, always in [1, K]
    S = 2 * (L + gap + HOP) // CHUNK                 # (Lpad - P) / K chunks from each of 2 views
    if S + c > PE_MAX or CHUNK + c > PE_MAX:
        raise ValueError(f"sequence of {S + c} tokens exceeds the positional table ({PE_MAX})")
    T_est = ENC_S * (L - 1) + ENC_K
    return PathShape(B=B, T=T, c=c, spk=spk, L=L, gap=gap, S=S, T_est=T_est)


def algorithmic_flops(ps: PathShape) -> float:
    """multiply-add = 2 FLOP; softmax/norm/activation not counted."""
    N, F, K = N_CH, D_FFN, CHUNK
    tok_i = ps.B * ps.S * (K + ps.c)
    tok_e = ps.B * K * (ps.S + ps.c)
    p0 = 2 * (3 * N * N + N * N + 2 * N * F)
    f_tr = N_LAYER * N_BLOCK * (tok_i * (p0 + 4 * (K + ps.c) * N) + tok_e * (p0 + 4 * (ps.S + ps.c) * N))
    f_other = 2 * ps.B * (ps.L * N * N + K * ps.S * N * N * ps.spk + 3 * ps.spk * ps.L * N * N
                          + ENC_K * N * ps.L * (1 + ps.spk)) + 8 * ps.B * ps.c * CTX_DIM * N
    return float(f_tr + f_other)
