"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

import cse_b200  # noqa: F401
from cse_b200 import synth
from cases import LOSS_CASES, MODEL_CASES  # noqa: F401

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        # (string-typed entries, e.g. the parameter-name lists of the gradient fixtures, stay numpy arrays)
        return {k: (torch.from_numpy(z[k]) if z[k].dtype.kind in "fiub" else z[k]) for k in z.files}


def model_case(name):
    """(state_dict, mix, sources, ctx, se, meta) for a MODEL_CASES entry — same generators as
    tests/golden/make_golden.py."""
    variant, spk, ce, c, B, T, cue, wseed, iseed = MODEL_CASES[name]
    sd = synth.make_state_dict(variant, spk, seed=wseed, ce=ce)
    mix, src = synth.make_mixture(B, T, max(spk, 2), seed=iseed)
    ctx = synth.make_context(B, c, seed=iseed) if variant != "sepformer" else None
    se = synth.make_speaker_embedding(B, seed=iseed) if variant == "hcontext" else None
    meta = dict(variant=variant, spk=spk, ce=ce, c=c, B=B, T=T, cue=cue)
    return sd, mix, src, ctx, se, meta


def loss_case(name):
    kind, B, T, C, seed = LOSS_CASES[name]
    _, a = synth.make_mixture(B, T, max(C, 2), seed=seed)
    _, b = synth.make_mixture(B, T, max(C, 2), seed=seed + 1000)
    a, b = a[:, :, :C], b[:, :, :C]
    est = 0.6 * a + 0.4 * b.flip(-1)
    return kind, est.contiguous(), a.contiguous()
