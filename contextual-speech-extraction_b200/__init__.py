"""B200-native Sepformer hot path for contextual speech extraction.

    import cse_b200
    from cse_b200.models.ContSep import Sepformer
"""
from . import shapes, synth  # noqa: F401
from . import _lib, runtime, modules, losses, backward, models  # noqa: F401
from .models import ContExt, ContSep, CSE_transformer, sepformer  # noqa: F401

__all__ = ["shapes", "synth", "modules", "losses", "backward", "models"]
