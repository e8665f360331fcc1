"""Import the reference's own model code, UNCHANGED, over the speechbrain shim.
TEST INFRASTRUCTURE ONLY; works only where /root/reference exists (the authoring container).
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("CSE_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sb_shim")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "models", "ContSep.py"))


def _paths():
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)


def build_reference_model(variant="contsep", num_spks=2, ce=True):
    """Instantiate the reference module for `variant` exactly as its train scripts do
    (train_ContSep.py:170-214, train_ContExt.py, train_HContExt.py)."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _paths()
    if variant == "sepformer":
        from src.models.sepformer import Sepformer
        return Sepformer(num_spks=num_spks)
    if variant == "contsep":
        from src.models.ContSep import Sepformer
        m = Sepformer(num_spks=num_spks, add_mt=True, ce=ce)
        m.add_mt_pipeline()
        return m
    from src.models.ContExt import Sepformer
    if variant == "context":
        m = Sepformer(num_spks=num_spks, add_ctx=True)
        m.add_ctx_pipeline()
        return m
    if variant == "hcontext":
        m = Sepformer(num_spks=num_spks, add_ctx=True, add_se=True)
        m.add_ctx_pipeline()
        m.add_se_pipeline()
        return m
    raise ValueError(variant)


def reference_losses():
    _paths()
    from speechbrain.nnet import losses
    import torchmetrics
    return losses.cal_si_snr, losses.get_si_snr_with_pitwrapper, torchmetrics.audio.ScaleInvariantSignalNoiseRatio
