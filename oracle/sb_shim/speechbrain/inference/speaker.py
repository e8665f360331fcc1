"""Name-only stub for train_HContExt.py:29 (ECAPA front-end is out of scope)."""


class EncoderClassifier:  # pragma: no cover
    @classmethod
    def from_hparams(cls, *a, **k):
        raise NotImplementedError("pretrained ECAPA is outside the hot path")
