"""B200-native Sepformer hot path for contextual speech extraction."""
from . import shapes, synth  # noqa: F401
