"""Drop-in mirror of the reference's `src/models` package (same module and class names):
    from cse_b200.models.ContSep import Sepformer        # src/models/ContSep.py
    from cse_b200.models.ContExt import Sepformer        # src/models/ContExt.py
    from cse_b200.models.sepformer import Sepformer      # src/models/sepformer.py
    from cse_b200.models.CSE_transformer import SBTransformerBlock_CSE
"""
