"""Numerical model of the planned bf16 tensor-core training step.  TEST INFRASTRUCTURE ONLY.

Inside `emulate_bf16_gemms()` every contraction the oracle performs (`@`, `torch.matmul`,
`torch.einsum`) behaves like a tensor-core GEMM of the CSE_BF16 mode in BOTH directions:
operands rounded to bfloat16, products accumulated at the tensor's working precision, and — in the
backward pass — the incoming gradient rounded to bfloat16 before the dgrad / wgrad contractions,
which reuse the rounded forward operands.  Everything else (residual stream, LayerNorm, GroupNorm,
softmax, losses) keeps full precision, as in DESIGN.md §3.

Purpose: decide on CPU, before any kernel exists, whether that design's gradients stay inside the
drift of the reference's OWN bf16-autocast training step (recorded in tests/golden/grad_*.npz by
make_golden_grads.py) — see tests/test_backward_oracle.py::test_bf16_backward_design_within_reference_drift.
"""
import contextlib

import torch
from torch.overrides import TorchFunctionMode


def _round_bf16(x):
    return x.to(torch.bfloat16).to(x.dtype)


class _RoundForward(torch.autograd.Function):
    """Round to bf16 going forward, pass the gradient through unchanged."""

    @staticmethod
    def forward(ctx, x):
        return _round_bf16(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBackward(torch.autograd.Function):
    """Identity going forward, round the gradient to bf16 on the way back."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _round_bf16(g)


_CONTRACTIONS = {torch.matmul, torch.Tensor.matmul, torch.Tensor.__matmul__, torch.mm, torch.bmm,
                 torch.Tensor.mm, torch.Tensor.bmm}


class _Bf16GemmMode(TorchFunctionMode):
    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in _CONTRACTIONS:
            a, b = args[0], args[1]
            with torch._C.DisableTorchFunction():
                return _RoundBackward.apply(torch.matmul(_RoundForward.apply(a), _RoundForward.apply(b)))
        if func is torch.einsum:
            eq, ops = args[0], args[1:]
            if len(ops) == 1 and isinstance(ops[0], (list, tuple)):
                ops = tuple(ops[0])
            with torch._C.DisableTorchFunction():
                return _RoundBackward.apply(torch.einsum(eq, *[_RoundForward.apply(o) for o in ops]))
        return func(*args, **kwargs)


@contextlib.contextmanager
def emulate_bf16_gemms():
    with _Bf16GemmMode():
        yield
