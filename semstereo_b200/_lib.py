"""ctypes binding of libsemstereo_b200.so (the C-ABI declared in include/semstereo_b200.h).

There is deliberately no fallback: if the library cannot be loaded, or a call is made without a CUDA
device, an exception is raised.  The product path never routes through torch/cuDNN/CPU re-implementations.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsemstereo_b200.so")

_P = ctypes.c_void_p
_I = ctypes.c_int
_F = ctypes.c_float

# name -> argtypes (all return int unless listed in _RESTYPES)
SIGNATURES = {
    "ss_version": [],
    "ss_last_error": [],
    "ss_sm_count": [],
    "ss_gwc_volume": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "ss_concat_volume": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ss_patch_gate": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ss_pointwise_conv2d": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ss_conv3d_f32": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "ss_conv3d_cout1_f32": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ss_to_blocked_bf16": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ss_widen_bf16": [_P, _P, ctypes.c_longlong, _P],
    "ss_from_blocked_bf16": [_P, _P, _I, _I, _I, _I, _I, _P],
    "ss_blocked_to_s2d": [_P, _P, _I, _I, _I, _I, _I, _P],
    "ss_window_attention_core_blocked": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "ss_gate_sigmoid_blocked": [_P, _P, _I, _I, _I, _I, _P],
    "ss_patch_gate_blocked": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ss_sparse_concat_volume_blocked": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ss_conv3d_tc_ntile": [_I, _I, _I],
    "ss_conv3d_tc_head": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ss_concat_stem_fused": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "ss_gwc_volume_backward": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "ss_concat_volume_backward": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ss_disparity_regression_backward": [_P, _P, _I, _I, _I, _I, _F, _P],
    "ss_regression_topk_backward": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ss_context_upsample_backward": [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    "ss_propagation_backward": [_P, _P, _I, _I, _I, _I, _P],
    "ss_disparity_variance_backward": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "ss_spatial_transformer_grid_backward": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ss_window_attention_core_f32_out": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "ss_window_attention_core_f32_masked": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "ss_window_attention_core_backward": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "ss_window_attention_core_backward_masked": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "ss_bilinear_up4": [_P, _P, _I, _I, _I, _P],
    "ss_bilinear_up4_backward": [_P, _P, _I, _I, _I, _P],
    "ss_conv3d_wgrad_f32": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "ss_conv2d_small_f32": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ss_conv2d_small_wgrad_supported": [_I, _I, _I],
    "ss_conv2d_small_wgrad_f32": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ss_bn_workspace_bytes": [_I],
    "ss_bn_train_forward": [_P, _P, _P, _P, _P, _P, _P, _I, _I, ctypes.c_longlong, _F, _I, _P],
    "ss_bn_train_backward": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, ctypes.c_longlong, _F, _P],
    "ss_conv2d_tc_ntile": [_I, _I, _I],
    "ss_conv2d_tc": [_I, _P, _I, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ss_conv2d_tc_ex": [_I, _P, _I, _P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ss_stem_conv3x3_s2": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "ss_dwconv3x3_blocked": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ss_groupnorm1_workspace_floats": [_I],
    "ss_groupnorm1_blocked": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "ss_linear_attention_workspace_floats": [_I, _I],
    "ss_linear_attention_blocked": [_P, _P, _P, _I, _I, _I, _I, _P],
    "ss_bilinear_up2": [_P, _P, _I, _I, _I, _P],
    "ss_pointwise_blocked_small": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ss_conv3d_tc": [_I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "ss_conv3d_tc_ex": [_I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "ss_conv3d_tc_split_supported": [_I, _I, _I],
    "ss_to_blocked_bf16_ex": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "ss_patch_gate_blocked_ex": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ss_conv3d_tc_head_ex": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ss_window_attention_core_f32": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "ss_window_attention3d": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "ss_att_stats": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "ss_sample_strength": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "ss_topk_select": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    "ss_sparse_concat_volume": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ss_regression_topk": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ss_ssr_param_count": [_I],
    "ss_ssr_upsample": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "ss_ssr_upsample2": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "ss_context_upsample": [_P, _P, _P, _I, _I, _I, _P],
    "ss_disparity_regression": [_P, _P, _I, _I, _I, _I, _F, _P],
    "ss_disparity_variance": [_P, _P, _P, _I, _I, _I, _I, _F, _P],
    "ss_propagation": [_P, _P, _I, _I, _I, _I, _P],
    "ss_spatial_transformer_grid": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
}
_RESTYPES = {"ss_last_error": ctypes.c_char_p}

_lib = None
_lock = threading.Lock()


class SemStereoLibraryError(RuntimeError):
    pass


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Loads (building in-tree with nvcc first if needed) the CUDA library.  Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if build_if_missing:
            from . import build as _build
            try:
                _build.build()               # no-op when the in-tree .so matches the sources
            except Exception as e:           # a stale .so is never silently used
                raise SemStereoLibraryError(f"cannot build {LIB_PATH}: {e}") from e
        if not os.path.exists(LIB_PATH):
            raise SemStereoLibraryError(f"{LIB_PATH} is missing; run `python -m semstereo_b200.build`")
        lib = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, ctypes.c_int)
        _lib = lib
    return _lib


def last_error() -> str:
    return (load().ss_last_error() or b"").decode()


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    msg = last_error() or what
    if rc == -1:
        raise ValueError(msg)
    if rc == -2:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)
