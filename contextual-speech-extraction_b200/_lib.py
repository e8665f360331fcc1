"""ctypes binding of csrc/libcse_b200.so (the C ABI declared in include/cse_b200.h).

The product path has no CPU fallback: if the shared library is missing or fails to load, every
call raises.  PyTorch is used only for device memory, streams and autograd plumbing.
"""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libcse_b200.so")

FP32, BF16 = 0, 1
N, K_CHUNK, LAYERS, BLOCKS, FFN, HEADS, CTX = 256, 250, 8, 2, 1024, 8, 4096

_f = C.POINTER(C.c_float)
_v = C.c_void_p


class LayerParams(C.Structure):
    _fields_ = [(n, _v) for n in (
        "in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b", "ffn1_w", "ffn1_b", "ffn2_w", "ffn2_b",
        "ln1_g", "ln1_b", "ln2_g", "ln2_b",
        "in_proj_w_bf16", "out_proj_w_bf16", "ffn1_w_bf16", "ffn2_w_bf16")]


class LayerGrads(C.Structure):
    """cse_layer_grads: gradient buffers of one transformer layer (accumulated into)."""
    _fields_ = [(n, _v) for n in (
        "in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b", "ffn1_w", "ffn1_b", "ffn2_w", "ffn2_b",
        "ln1_g", "ln1_b", "ln2_g", "ln2_b")]


class StackParams(C.Structure):
    _fields_ = [("layer", LayerParams * LAYERS), ("final_g", _v), ("final_b", _v), ("pe", _v)]


class BlockParams(C.Structure):
    _fields_ = [("intra", StackParams), ("inter", StackParams)] + [(n, _v) for n in (
        "intra_norm_g", "intra_norm_b", "inter_norm_g", "inter_norm_b",
        "intra_map_w", "intra_map_b", "inter_map_w", "inter_map_b")]


class Params(C.Structure):
    _fields_ = ([(n, _v) for n in ("enc_w", "norm_g", "norm_b", "conv1d_w")]
                + [("block", BlockParams * BLOCKS)]
                + [(n, _v) for n in ("prelu", "conv2d_w", "conv2d_b", "out_w", "out_b", "gate_w", "gate_b",
                                     "end_w", "dec_w", "conv1d_w_bf16", "conv2d_w_bf16", "out_w_bf16",
                                     "gate_w_bf16", "end_w_bf16")])


class Shape(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("B", "T", "c", "spk", "L", "gap", "S", "T_est")]


_SIGNATURES = {
    "cse_version": (C.c_int, []),
    "cse_last_error": (C.c_char_p, []),
    "cse_launch_count": (C.c_longlong, []),
    "cse_debug_force_mma_attention": (C.c_int, [C.c_int]),
    "cse_debug_attention_trace": (C.c_int, [_v]),
    "cse_profile_enable": (C.c_int, [C.c_int]),
    "cse_profile_collect": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int]),
    "cse_path_shape": (C.c_int, [C.c_int] * 4 + [C.POINTER(Shape)]),
    "cse_workspace_bytes": (C.c_size_t, [C.c_int] * 5),
    "cse_pack_bf16_elems": (C.c_size_t, [C.c_int]),
    "cse_pack_bf16": (C.c_int, [C.POINTER(Params), C.c_int, _v, C.c_size_t, _v]),
    "cse_forward": (C.c_int, [C.POINTER(Params), _v, _v] + [C.c_int] * 5 + [_v, _v, _v, C.c_size_t, _v]),
    "cse_forward_host": (C.c_int, [C.POINTER(Params), _v, _v] + [C.c_int] * 5 + [_v, _v, _v, C.c_size_t, _v]),
    "cse_pipeline_workspace_bytes": (C.c_size_t, [C.c_int] * 6),
    "cse_pipeline_create": (C.c_int, [C.POINTER(Params)] + [C.c_int] * 6 + [_v, C.c_size_t, C.POINTER(_v)]),
    "cse_pipeline_submit": (C.c_int, [_v, _v, _v, _v, _v, C.POINTER(C.c_int)]),
    "cse_pipeline_wait": (C.c_int, [_v, C.c_int]),
    "cse_pipeline_destroy": (C.c_int, [_v]),
    "cse_masknet_fwd": (C.c_int, [C.POINTER(Params), _v, _v] + [C.c_int] * 5 + [_v, _v, _v, C.c_size_t, _v]),
    "cse_encoder_fwd": (C.c_int, [_v, _v, C.c_int, C.c_int, C.c_int, _v, _v, C.POINTER(C.c_int), _v]),
    "cse_gn_finalize": (C.c_int, [_v, C.c_int, C.c_int, C.c_double, C.c_float, _v, _v]),
    "cse_gn_apply": (C.c_int, [_v, _v, _v, _v, C.c_int, C.c_int, C.c_int, _v, _v]),
    "cse_linear": (C.c_int, [_v, C.c_int, _v, _v, C.c_float, _v, _v, C.c_int, C.c_int, C.c_int, C.c_int,
                             C.c_int, C.c_int, C.c_int, _v]),
    "cse_layernorm_fwd": (C.c_int, [_v, _v, _v, C.c_int, C.c_float, C.c_int, _v, _v]),
    "cse_ln_linear": (C.c_int, [_v, _v, _v, C.c_float, _v, _v, _v, C.c_int, C.c_int, C.c_int, C.c_int, _v]),
    "cse_ffn_ln_fused": (C.c_int, [_v, _v, _v, C.c_float, _v, _v, _v, _v, _v, _v, _v, _v, C.c_int, _v]),
    "cse_linear_residual_ln": (C.c_int, [_v, C.c_int, _v, _v, _v, _v, _v, C.c_float, _v, C.c_int, C.c_int, _v]),
    "cse_ffn_fused": (C.c_int, [_v, _v, _v, _v, _v, _v, C.c_int, _v]),
    "cse_attention_fwd": (C.c_int, [_v, C.c_int, C.c_int, C.c_int, _v, _v]),
    "cse_segment": (C.c_int, [_v, C.c_int, C.c_int, C.c_int, _v, _v]),
    "cse_build_sequences": (C.c_int, [_v, _v, _v, C.c_int, C.c_int, C.c_int, C.c_int, _v, _v]),
    "cse_context_map": (C.c_int, [_v, _v, _v, C.c_int, C.c_int, _v, _v]),
    "cse_stack_finish": (C.c_int, [_v] * 6 + [C.c_int] * 4 + [_v, _v, _v, _v]),
    "cse_pred_head": (C.c_int, [_v, _v, _v, C.c_int, C.c_int, C.c_int, _v, _v]),
    "cse_prelu_overlap_add": (C.c_int, [_v, _v, C.c_int, C.c_int, C.c_int, C.c_int, _v, _v]),
    "cse_gate": (C.c_int, [_v, _v, C.c_size_t, C.c_int, _v, _v]),
    "cse_mask_decode": (C.c_int, [_v, _v, _v, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _v, _v, _v]),
    "cse_si_snr": (C.c_int, [_v, _v, C.c_int, C.c_int, C.c_int, _v, _v]),
    "cse_pit_si_snr": (C.c_int, [_v, _v, C.c_int, C.c_int, C.c_int, _v, _v, _v]),
    "cse_tm_si_snr": (C.c_int, [_v, _v, C.c_int, C.c_int, _v, _v]),
    "cse_selection_loss": (C.c_int, [_v, _v, _v, C.c_int, C.c_int, C.c_int, C.c_int, _v, _v, _v, _v, _v, _v]),
    "cse_select_stream": (C.c_int, [_v, _v, C.c_int, C.c_int, C.c_int, C.c_int, _v, _v, _v]),
    "cse_selection_accuracy": (C.c_int, [_v, _v, C.c_int, C.c_int, C.c_int, _v, _v, _v]),
    "cse_si_snr_bwd": (C.c_int, [_v, _v, _v, C.c_int, C.c_int, C.c_int, _v, _v, _v]),
    "cse_pit_si_snr_bwd": (C.c_int, [_v, _v, _v, _v, C.c_int, C.c_int, C.c_int, _v, _v, _v]),
    "cse_tm_si_snr_bwd": (C.c_int, [_v, _v, _v, C.c_int, C.c_int, _v, _v, _v]),
    "cse_linear_bwd": (C.c_int, [_v, C.c_int, _v, _v, C.c_int, C.c_int, C.c_int, C.c_int, _v, C.c_int, _v, _v,
                                 _v, _v]),
    "cse_layernorm_bwd": (C.c_int, [_v, _v, _v, C.c_int, C.c_float, C.c_int, _v, _v, _v, _v]),
    "cse_attention_bwd": (C.c_int, [_v, _v, _v, C.c_int, C.c_int, _v, _v]),
    "cse_attention_bwd_bf16": (C.c_int, [_v, _v, _v, C.c_int, C.c_int, _v, _v]),
    "cse_groupnorm_fwd": (C.c_int, [_v, _v, _v, _v, C.c_int, C.c_int, C.c_float, _v, _v, _v, _v]),
    "cse_groupnorm_bwd": (C.c_int, [_v, _v, _v, _v, C.c_int, C.c_int, _v, _v, _v, _v, _v]),
    "cse_sequences_to_chunks": (C.c_int, [_v, C.c_int, C.c_int, C.c_int, C.c_int, _v, _v, _v]),
    "cse_prelu_overlap_add_bwd": (C.c_int, [_v, _v, _v, C.c_int, C.c_int, C.c_int, _v, _v, _v]),
    "cse_gate_bwd": (C.c_int, [_v, _v, _v, C.c_size_t, _v, _v, _v]),
    "cse_mask_decode_bwd": (C.c_int, [_v, _v, _v, _v, C.c_int, C.c_int, C.c_int, C.c_int, _v, _v, _v, _v]),
    "cse_encoder_bwd": (C.c_int, [_v, _v, _v, C.c_int, C.c_int, _v, _v]),
    "cse_linear_bwd_tc_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "cse_linear_bwd_tc": (C.c_int, [_v, C.c_int, C.c_int, _v, _v, C.c_int, C.c_int, C.c_int, _v, C.c_int, C.c_int,
                                    _v, _v, _v, C.c_size_t, _v]),
    "cse_layer_bwd_bf16_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "cse_layer_bwd_bf16": (C.c_int, [C.POINTER(LayerParams), C.POINTER(LayerGrads), _v, _v, C.c_int, C.c_int, _v,
                                     C.c_size_t, _v]),
    "cse_sdr_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "cse_sdr": (C.c_int, [_v, _v, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _v, _v, C.c_size_t, _v]),
    "cse_metric_update": (C.c_int, [_v, C.c_int, _v, _v]),
    "cse_mix_audio": (C.c_int, [_v] * 8 + [C.c_int, C.c_int, C.c_int, C.c_longlong] + [_v] * 6),
    "cse_peak_normalize": (C.c_int, [_v, _v, C.c_int, C.c_float, C.c_longlong, _v, _v]),
    "cse_decimate": (C.c_int, [_v, _v, C.c_int, C.c_longlong, C.c_int, _v, C.c_int, C.c_longlong, _v, _v, _v]),
    "cse_optim_chunk_count": (C.c_longlong, [C.c_int, C.POINTER(C.c_longlong)]),
    "cse_optim_table_fill": (C.c_int, [C.c_int, C.POINTER(C.c_longlong)] + [C.POINTER(_v)] * 5 + [_v, C.c_size_t]),
    "cse_optim_table_set_grads": (C.c_int, [C.c_int, C.POINTER(C.c_longlong), C.POINTER(_v), _v, C.c_size_t]),
    "cse_optim_step": (C.c_int, [_v, C.c_longlong] + [C.c_float] * 5 + [C.c_int, C.c_float, C.c_int, C.c_float,
                                 C.c_float, C.c_int, C.c_int, _v, _v, _v]),
    "cse_layer_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "cse_layer_fwd": (C.c_int, [C.POINTER(LayerParams), _v, C.c_int, C.c_int, C.c_int, _v, C.c_size_t, _v]),
    "cse_layer_bwd": (C.c_int, [C.POINTER(LayerParams), C.POINTER(LayerGrads), _v, _v, C.c_int, C.c_int, _v,
                                C.c_size_t, _v]),
}

_lock = threading.Lock()
_lib = None


class CseError(RuntimeError):
    """Raised when a C-ABI call returns non-zero (message from cse_last_error)."""


def load():
    """Load the shared library once; raise loudly if it is not built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.isfile(LIB_PATH):
                raise CseError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(there is no CPU fallback for the CUDA path)")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def exported_names():
    return list(_SIGNATURES)


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise CseError(f"{name}: {lib.cse_last_error().decode(errors='replace')}")


def path_shape(B, T, c, spk):
    s = Shape()
    call("cse_path_shape", B, T, c, spk, C.byref(s))
    return s


def ptr(t):
    """Device/host pointer of a tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
