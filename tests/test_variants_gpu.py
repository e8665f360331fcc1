"""GPU: the opt-in kernel variants (selected by environment variables that the library reads once per
process) produce the same results as the default path.  Each variant runs in a fresh interpreter:
the bf16 golden-vector tests of test_forward_gpu.py plus the tensor-core GEMM tests.

CSE_FFN_LN=0 / 1    separate layernorm_kernel launches / the LayerNorm warps of the feed-forward kernel at every size
                    (default: the latter from 16 k rows up, so the small golden fixtures need the switch to reach it)
CSE_FFN_FUSED=0     two-GEMM feed-forward instead of the fused kernel (ffn_tc.cu)
CSE_LN_FUSED=1      norm1 -> in_proj as one kernel (gemm_ln_tc.cu)
CSE_OUTPROJ_LN=1    out-proj + residual + norm2 as one kernel (gemm_tc.cu, LayerNorm epilogue)
CSE_ATTN_VER=4      the 16-softmax-warp tcgen05 attention kernel (attention_tc.cu)  [+ CSE_DECODE_SIMT=1: the SIMT
                    decoder contraction in bf16 mode (head.cu), in the same interpreter]
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{"CSE_FFN_LN": "0"}, {"CSE_FFN_LN": "1"}, {"CSE_FFN_FUSED": "0"}, {"CSE_LN_FUSED": "1"}, {"CSE_OUTPROJ_LN": "1"},
                                 {"CSE_ATTN_VER": "4", "CSE_DECODE_SIMT": "1"}],
                         ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_variant_matches_golden(env):
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
           os.path.join(ROOT, "tests", "test_forward_gpu.py"), "-k", "bf16 or graph",
           os.path.join(ROOT, "tests", "test_gemm_tc_gpu.py")]
    r = subprocess.run(cmd, cwd=ROOT, env={**os.environ, **env}, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_transposed_copy_backward_variant():
    """CSE_WGRAD_TRANSPOSE=1: dgrad / wgrad through transposed bf16 copies (the first tensor-core backward)."""
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
           os.path.join(ROOT, "tests", "test_backward_tc_gpu.py"), "-k", "linear or layer"]
    r = subprocess.run(cmd, cwd=ROOT, env={**os.environ, "CSE_WGRAD_TRANSPOSE": "1"}, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
