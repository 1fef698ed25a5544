"""ctypes binding of libzsg_b200.so (the C ABI declared in include/zsg_b200.h).

There is no fallback: if the shared library is missing or a call fails, this module raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzsg_b200.so")


class ZsgError(RuntimeError):
    pass


class RowT(C.Structure):
    _fields_ = [("base", C.c_int32), ("y0", C.c_int16), ("x0", C.c_int16), ("hin", C.c_int16), ("win", C.c_int16),
                ("out", C.c_int32)]


class ConvParams(C.Structure):
    _fields_ = [("x", C.c_void_p), ("w", C.c_void_p), ("w_lo", C.c_void_p), ("y", C.c_void_p), ("rows", C.c_void_p),
                ("m", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32), ("r", C.c_int32), ("s", C.c_int32),
                ("in_div", C.c_int32), ("in_scale", C.c_void_p), ("in_shift", C.c_void_p), ("in_relu", C.c_int32),
                ("bias", C.c_void_p), ("out_relu", C.c_int32), ("out_mask", C.c_void_p), ("residual", C.c_void_p), ("accumulate", C.c_int32),
                ("impl", C.c_int32), ("x_lo", C.c_void_p), ("dil", C.c_int32), ("stats", C.c_void_p),
                ("x_bf16", C.c_void_p), ("w_bf16", C.c_void_p), ("x_plain", C.c_int32), ("y_bf16", C.c_void_p), ("residual_bf16", C.c_void_p), ("y_pitch", C.c_int32), ("row_add", C.c_void_p),
                ("row_add_idx", C.c_void_p), ("y_lo", C.c_void_p), ("y_img_bf16", C.c_void_p),
                ("bnb_x", C.c_void_p), ("bnb_scale", C.c_void_p), ("bnb_shift", C.c_void_p), ("bnb_partials", C.c_void_p)]


class WgradParams(C.Structure):
    _fields_ = [("x", C.c_void_p), ("dy", C.c_void_p), ("dw", C.c_void_p), ("rows", C.c_void_p),
                ("m", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32), ("r", C.c_int32), ("s", C.c_int32),
                ("in_scale", C.c_void_p), ("in_shift", C.c_void_p), ("in_relu", C.c_int32), ("split_k", C.c_int32),
                ("impl", C.c_int32), ("x_lo", C.c_void_p), ("dy_lo", C.c_void_p), ("dy_pitch", C.c_int32), ("dil", C.c_int32),
                ("x_bf16", C.c_void_p), ("dy_bf16", C.c_void_p)]


_P, _I, _L, _F, _D, _Z = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t

# name -> argtypes (return type is int unless listed in _RET)
SIGNATURES = {
    "zsg_last_error_string": [],
    "zsg_abi_version": [],
    "zsg_device_supported": [],
    "zsg_conv_fwd": [C.POINTER(ConvParams), _P],
    "zsg_conv_wgrad": [C.POINTER(WgradParams), _P],
    "zsg_debug_set_conv_trace": [_P, _I],
    "zsg_weight_transpose_flip": [_P, _P, _I, _I, _I, _I, _P],
    "zsg_weight_transpose_flip_batched": [_P, _P, _P, _I, _L, _P],
    "zsg_weight_transpose_flip_batched32": [_P, _P, _P, _I, _L, _P],
    "zsg_pad_channels": [_P, _P, _L, _I, _I, _P],
    "zsg_split_tf32": [_P, _P, _P, _L, _P],
    "zsg_split_act": [_P, _P, _P, _I, _P, _P, _L, _I, _P],
    "zsg_cast_bf16": [_P, _P, _P, _I, _P, _L, _I, _P],
    "zsg_act_b16": [_P, _P, _P, _I, _P, _L, _I, _P],
    "zsg_bn_apply_b16": [_P, _P, _P, _P, _P, _P, _I, _P, _L, _I, _P],
    "zsg_bn_bwd_reduce_b16": [_P, _I, _P, _P, _P, _P, _P, _P, _I, _P, _I, _P, _L, _I, _P],
    "zsg_bn_bwd_apply_b16": [_P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _L, _I, _P],
    "zsg_maxpool_bn_relu_fwd_b16": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "zsg_bn_apply_bf16": [_P, _P, _P, _P, _P, _P, _I, _P, _P, _L, _I, _P],
    "zsg_bn_bwd_apply_bf16": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _L, _I, _P],
    "zsg_nchw_to_nhwc4": [_P, _P, _I, _I, _I, _P],
    "zsg_nchw_to_nhwc8_bf16": [_P, _P, _I, _I, _I, _P],
    "zsg_colsum": [_P, _P, _L, _I, _I, _I, _P],
    "zsg_gather_rows": [_P, _P, _P, _L, _I, _I, _P],
    "zsg_bn_stats": [_P, _P, _L, _I, _P],
    "zsg_bn_stats_partials": [_P, _L, _I, _P, _P],
    "zsg_bn_bwd_center_sums": [_P, _P, _P, _I, _P],
    "zsg_bn_finalize_partials": [_P, _L, _L, _I, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "zsg_bn_finalize": [_P, _L, _I, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P],
    "zsg_bn_eval_affine": [_P, _P, _P, _P, _F, _I, _P, _P, _P],
    "zsg_bn_apply": [_P, _P, _P, _P, _P, _P, _I, _P, _P, _L, _I, _P],
    "zsg_bn_bwd_reduce": [_P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _L, _I, _P],
    "zsg_bn_bwd_apply": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _L, _I, _P],
    "zsg_maxpool_bn_relu_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "zsg_maxpool_bn_relu_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "zsg_upsample_add": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "zsg_upsample_add_bwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "zsg_avgpool_fwd": [_P, _P, _I, _I, _I, _P],
    "zsg_avgpool_bwd": [_P, _P, _I, _I, _I, _P],
    "zsg_maxpool_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "zsg_maxpool_bwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "zsg_l2norm_fwd": [_P, _P, _P, _L, _I, _P],
    "zsg_l2norm_bwd": [_P, _P, _P, _P, _L, _I, _I, _I, _P],
    "zsg_relu_bwd": [_P, _P, _P, _L, _I, _P],
    "zsg_axpy": [_P, _P, _F, _L, _P],
    "zsg_scale_dev": [_P, _L, _I, _L, _P, _P],
    "zsg_copy_cols": [_P, _L, _P, _L, _L, _I, _P],
    "zsg_head0_lang_grid_terms": [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    "zsg_head0_backward_sums": [_P, _P, _P, _P, _P, _I, _I, _I, _P, _Z, _P, _P, _L, _P],
    "zsg_fuse_lang_grid": [_P, _P, _P, _P, _I, _I, C.POINTER(C.c_int32), _I, _I, _I, _I, _P],
    "zsg_unfuse_lang_grid": [_P, _P, _P, _I, _I, C.POINTER(C.c_int32), _I, _I, _I, _I, _P],
    "zsg_lstm_fwd_dir": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P],
    "zsg_lstm_rev_step": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P],
    "zsg_lstm_bwd_dir": [_P, _P, _P, _P, _P, _P, _I, _I, _P, _P],
    "zsg_lstm_rev_step_bwd": [_P, _P, _P, _I, _P, _P],
    "zsg_match_loss_workspace_bytes": [_I],
    "zsg_match_loss": [_P, _L, _P, _L, _P, _P, _I, _I, _D, _F, _F, _D, _I, _P, _P, _L, _P, _L, _P, _P, _P, _Z, _P],
    "zsg_match": [_P, _P, _I, _I, _D, _I, _P, _P, _P, _Z, _P],
    "zsg_loss_grad": [_P, _L, _P, _L, _P, _P, _P, _I, _I, _F, _F, _D, _P, _P, _L, _P, _L, _P, _Z, _P],
    "zsg_eval_workspace_bytes": [_I],
    "zsg_eval": [_P, _L, _P, _L, _P, _P, _P, _I, _I, _D, _P, _P, _P, _P, _P, _Z, _P],
    "zsg_adam": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _I, _F, _P],
    "zsg_resize_rgb8": [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P],
    "zsg_embed_gather": [_P, _P, _P, _L, _I, _P],
}
_RET = {"zsg_last_error_string": C.c_char_p, "zsg_match_loss_workspace_bytes": C.c_size_t,
        "zsg_eval_workspace_bytes": C.c_size_t}

_lib = None


def load():
    """Load the shared library; raises ZsgError if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ZsgError(f"{LIB_PATH} not found: build it with `python zsgnet-pytorch_b200/build.py` "
                       "(the CUDA extension is mandatory; there is no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.argtypes = argtypes
        fn.restype = _RET.get(name, C.c_int)
    _lib = lib
    return lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


# kernels launched per C-ABI call (memsets not counted); used for bench.py's gpu_launches claim
KERNELS_PER_CALL = {"zsg_match_loss": 3, "zsg_loss_grad": 2, "zsg_unfuse_lang_grid": 2, "zsg_resize_rgb8": 2,
                    "zsg_head0_lang_grid_terms": 2, "zsg_head0_backward_sums": 2}
LAUNCH_COUNT = [0]


def call(name, *args):
    lib = load()
    LAUNCH_COUNT[0] += KERNELS_PER_CALL.get(name, 1)
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise ZsgError(f"{name} failed ({rc}): {lib.zsg_last_error_string().decode()}")
    return rc
