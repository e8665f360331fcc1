"""Mirror of src/models/CSE_transformer.py (class names and constructor signatures)."""
from ..modules import (MultiheadAttention, SBTransformerBlock_CSE, TransformerEncoder,  # noqa: F401
                       TransformerEncoderLayer)
