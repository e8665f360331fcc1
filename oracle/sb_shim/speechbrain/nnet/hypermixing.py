"""Name-only stub (CSE_transformer.py:325-332 branch unused)."""


class HyperMixing:  # pragma: no cover
    def __init__(self, *a, **k):
        raise NotImplementedError("HyperMixing is outside the hot path")
