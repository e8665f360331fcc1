"""Golden vectors for the loader-side mixture synthesis, produced by the reference's OWN functions
(/root/reference/mix_aud.py — pure numpy, identical to the methods of dataset_train_CSE.py:417-505).

    python tests/golden/make_golden_mixture.py          # needs /root/reference; writes tests/golden/mixture_*.npz

Inputs are regenerated from the seeds by tests/test_mixture.py (`mixture_case`), so only outputs are stored (as
float32, which is what the reference's collate_fn keeps)."""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from cases import MIXTURE_CASES, mixture_case  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location("ref_mix_aud", "/root/reference/mix_aud.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for name in MIXTURE_CASES:
        clips, snrs, pad = mixture_case(name)
        if len(clips) == 2:
            outs = ref.mix_audio(clips[0], clips[1], snrs[0], pad=pad)
        else:
            outs = ref.mix_audio_3spk(clips[0], clips[1], clips[2], snrs[0], snrs[1], pad=pad)
        assert all(o.dtype == np.float64 for o in outs), [o.dtype for o in outs]
        np.savez_compressed(os.path.join(HERE, f"mixture_{name}.npz"),
                            **{f"out{i}": o.astype(np.float32) for i, o in enumerate(outs)})
        print(name, [o.shape for o in outs])


if __name__ == "__main__":
    main()
