"""Minimal torchmetrics stand-in (test infrastructure): SI-SNR only (train_ContExt.py:339)."""
from . import audio  # noqa: F401
