"""Minimal speechbrain stand-in (test infrastructure; see ../README.md)."""
from . import nnet  # noqa: F401
from . import lobes  # noqa: F401
