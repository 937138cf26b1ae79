"""ctypes binding of libumhs_sm100.so (C ABI declared in include/umhs.h).

There is NO fallback: if the library is missing or an entry point fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint8, c_uint64, c_void_p
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# UMHS_LIB_PATH: developer override (A/B runs of kernel variants built by tools/build_variant.py)
LIB_PATH = os.environ.get("UMHS_LIB_PATH") or os.path.join(_HERE, "libumhs_sm100.so")

P = c_void_p  # every device pointer crosses the ABI as a plain address


class UmhsError(RuntimeError):
    pass


class Points(Structure):
    _fields_ = [("positions", P), ("directions", P), ("rays_o", P), ("rays_d", P), ("ray_indices", P),
                ("t_starts", P), ("t_ends", P), ("contraction", c_int32), ("aabb", c_float * 6), ("enc_index", P), ("n_dev", P)]


class HashGrid(Structure):
    _fields_ = [("table", P), ("scalings", P), ("num_levels", c_int32), ("log2_table_size", c_int32)]


_MLP_NAMES = ["base_w0", "base_b0", "base_w1", "base_b1", "head_w0", "head_b0", "head_w1", "head_b1", "head_w2",
              "head_b2", "feat_w0", "feat_b0", "feat_w1", "feat_b1", "feat_w2", "feat_b2", "dir_w0", "dir_b0"]


class MlpParams(Structure):
    _fields_ = [(n, P) for n in _MLP_NAMES] + [("num_classes", c_int32), ("specular", c_int32),
                                                 ("temperature", c_float), ("enc_tiles", c_int32)]


class MlpGrads(Structure):
    _fields_ = [(n, P) for n in _MLP_NAMES]


class FieldOut(Structure):
    _fields_ = [("sigma", P), ("embedding", P), ("scalar", P), ("abund", P), ("s1", P), ("spec_hidden", P),
                ("act_tiles", P)]


class FieldSaved(Structure):
    _fields_ = [("act_tiles", P), ("sigma", P), ("scalar", P), ("abund", P), ("s1", P), ("spec_hidden", P)]


class FieldOutGrads(Structure):
    _fields_ = [("d_sigma", P), ("d_embedding", P), ("d_scalar", P), ("d_abund", P), ("d_s1", P),
                ("d_spec_hidden", P)]


class CompositeIn(Structure):
    _fields_ = [("n_rays", c_int64), ("n_samples", c_int64), ("num_classes", c_int32), ("num_bands", c_int32),
                ("specular", c_int32), ("grad_scaling", c_int32), ("packed_info", P), ("t_starts", P), ("t_ends", P),
                ("sigma", P), ("scalar", P), ("abund", P), ("s1", P), ("spec_hidden", P), ("endmembers", P),
                ("dir_w1", P), ("dir_b1", P), ("rgb_matrix", P), ("palette", P), ("specular_in", P)]


class CompositeOut(Structure):
    _fields_ = [("trans", P), ("weights", P), ("spectral", P), ("spectral2", P), ("specular", P), ("abundances", P),
                ("accumulation", P), ("depth", P), ("t_minmax", P), ("rgb", P), ("seg_probs", P), ("seg_raw", P),
                ("seg_pred", P)]


class CompositeGradsIn(Structure):
    _fields_ = [("d_spectral", P), ("d_spectral2", P), ("d_rgb", P), ("d_accumulation", P), ("spectral", P),
                ("trans", P), ("d_weights", P)]


class CompositeGradsOut(Structure):
    _fields_ = [("d_sigma", P), ("d_scalar", P), ("d_abund", P), ("d_s1", P), ("d_spec_hidden", P),
                ("d_endmembers", P), ("d_dir_w1", P), ("d_dir_b1", P)]


class BatchIn(Structure):
    _fields_ = [("u", P), ("n_rays", c_int64), ("num_images", c_int32), ("height", c_int32), ("width", c_int32),
                ("num_bands", c_int32), ("hs_dtype", c_int32), ("image_channels", c_int32), ("image_dtype", c_int32),
                ("image_idx", P), ("hs", P), ("image", P), ("c2w", P), ("intrinsics", P)]


class BatchOut(Structure):
    _fields_ = [("indices", P), ("hs_image", P), ("image", P), ("origins", P), ("directions", P), ("pixel_area", P),
                ("directions_norm", P)]


class MarchArgs(Structure):
    _fields_ = [("rays_o", P), ("rays_d", P), ("near_planes", P), ("far_planes", P), ("binaries", P), ("aabbs", P),
                ("levels", c_int32), ("resolution", c_int32), ("step_size", c_float), ("cone_angle", c_float),
                ("n_rays", c_int64), ("level_any", P), ("occ_bits", P), ("stage", P), ("stage_counts", P), ("stage_stride", c_int32)]


# name -> argtypes (restype is int unless listed in _RESTYPES)
_SIGNATURES = {
    "umhs_last_error": [],
    "umhs_version": [],
    "umhs_hash_encode_fwd": [POINTER(Points), c_int64, POINTER(HashGrid), P, P],
    "umhs_hash_encode_bwd": [POINTER(Points), c_int64, POINTER(HashGrid), P, P, P],
    "umhs_hash_encode_fwd_tiles": [POINTER(Points), c_int64, POINTER(HashGrid), P, P],
    "umhs_hash_encode_bwd_tiles": [POINTER(Points), c_int64, POINTER(HashGrid), P, P, P],
    "umhs_hash_encode_bwd_tiles_levels": [POINTER(Points), c_int64, POINTER(HashGrid), P, P, c_int32, c_int32, P],
    "umhs_hash_encode_x01_fwd": [P, c_int64, POINTER(HashGrid), P, P],
    "umhs_hash_encode_x01_bwd": [P, c_int64, POINTER(HashGrid), P, P, P],
    "umhs_density_fwd": [POINTER(Points), c_int64, P, POINTER(MlpParams), P, P],
    "umhs_field_mlp_fwd": [POINTER(Points), c_int64, P, POINTER(MlpParams), POINTER(FieldOut), P],
    "umhs_field_mlp_bwd": [POINTER(Points), c_int64, P, POINTER(MlpParams), POINTER(FieldOutGrads), P,
                           POINTER(MlpGrads), P],
    "umhs_density_tc_fwd": [POINTER(Points), c_int64, P, POINTER(MlpParams), P, P],
    "umhs_field_mlp_tc_fwd": [POINTER(Points), c_int64, P, POINTER(MlpParams), POINTER(FieldOut), P],
    "umhs_field_mlp_tc_bwd": [POINTER(Points), c_int64, P, POINTER(MlpParams), POINTER(FieldOutGrads), P,
                              POINTER(MlpGrads), P],
    "umhs_field_mlp_tc_bwd_saved": [POINTER(Points), c_int64, P, POINTER(MlpParams), POINTER(FieldOutGrads),
                                    POINTER(FieldSaved), P, POINTER(MlpGrads), P],
    "umhs_field_mlp_tc_bwd_scatter": [POINTER(Points), c_int64, P, POINTER(MlpParams), POINTER(FieldOutGrads), P,
                                      POINTER(MlpGrads), POINTER(HashGrid), P, P],
    "umhs_debug_trace": [P, c_int32, P],
    "umhs_tc05_selftest": [c_int32, P, P, P, c_int32, c_int32, P],
    "umhs_tc05_microbench": [c_int32, c_int32, c_int32, c_int32, P, P],
    "umhs_composite_fwd": [POINTER(CompositeIn), POINTER(CompositeOut), P],
    "umhs_depth_clip": [P, c_int64, P, P],
    "umhs_depth_clip_chunks": [P, c_int64, P, P, P, P, c_int64, P, P],
    "umhs_composite_bwd": [POINTER(CompositeIn), POINTER(CompositeGradsIn), POINTER(CompositeGradsOut), P],
    "umhs_specular_tc_fwd": [c_int64, P, c_int64, c_int32, P, P, P, P, P, P, P, P],
    "umhs_specular_tc_bwd": [c_int64, P, c_int64, c_int32, c_int32, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P],
    "umhs_spectral_grad": [c_int64, c_int32, P, P, P, P, P, P],
    "umhs_unmix_fwd": [c_int64, c_int32, c_int32, c_int32, P, P, P, P, P, P, P, P, P, P],
    "umhs_unmix_bwd": [c_int64, c_int32, c_int32, c_int32, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P],
    "umhs_render_weights_fwd": [P, c_int64, P, P, P, P, P, P],
    "umhs_render_weights_bwd": [P, c_int64, P, P, P, P, P, P, P],
    "umhs_accumulate_fwd": [P, P, P, c_int64, c_int32, P, P],
    "umhs_accumulate_bwd": [P, P, P, c_int64, c_int32, P, P, P, P],
    "umhs_march_level_any": [POINTER(MarchArgs), P, P],
    "umhs_occ_pack_bits": [POINTER(MarchArgs), P, P, P],
    "umhs_march_count": [POINTER(MarchArgs), P, P],
    "umhs_pack_scan": [P, c_int64, P, P, c_int64, P],
    "umhs_march_write": [POINTER(MarchArgs), P, P, P, P, P],
    "umhs_visibility_mask": [P, c_int64, P, P, P, c_float, c_float, P, P, P, P],
    "umhs_compact_samples": [P, P, c_int64, P, P, P, P, P, P, P, P],
    "umhs_pack_info": [P, c_int64, c_int64, P, P, P],
    "umhs_occ_cell_points": [P, P, c_uint64, P, c_int64, c_int32, P, P, P],
    "umhs_occ_select_cells": [P, c_int64, c_int64, P, P, c_uint64, P, P, P, P, P],
    "umhs_occ_ema_update": [P, P, c_int64, P, c_float, c_float, c_int64, P, P],
    "umhs_occ_binarize": [P, c_int64, c_float, P, P, P, P],
    "umhs_adam_step": [P, P, P, P, c_int64, c_float, c_float, c_float, c_float, c_int32, P, P, c_int64, c_int64,
                       c_int32, P],
    "umhs_adam_step_dev": [P, P, P, P, c_int64, c_float, c_float, c_float, P, P, P, c_int64, c_int64, c_int32, P],
    "umhs_reduce_slabs": [P, c_int64, c_int64, c_int64, P, P],
    "umhs_sample_batch": [POINTER(BatchIn), POINTER(BatchOut), P],
    "umhs_train_loss": [c_int64, c_int32, c_int32, P, P, P, P, P, P, c_int32, c_float, c_float, P, P, P, P, P],
    "umhs_dp_barrier": [P, c_int32, c_int32, c_int32, P, P],
    "umhs_dp_adam_step": [P, P, P, P, c_int64, c_int32, c_int32, c_float, c_float, c_float, P, c_int64, c_int64, P, P,
                          P],
    "umhs_dp_adam_step_range": [P, P, P, P, c_int64, c_int64, c_int64, c_int32, c_int32, c_float, c_float, c_float, P, c_int64, c_int64, P, P,
                          P],
}
_RESTYPES = {"umhs_last_error": c_char_p}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib: Optional[ctypes.CDLL] = None


def load() -> ctypes.CDLL:
    """Load the library (once).  Raises UmhsError when it has not been built: there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UmhsError(f"{LIB_PATH} is missing: build it with `python -m umhs_b200.build` "
                        "(the UMHS B200 plugin has no CPU / PyTorch fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().umhs_last_error().decode()
        raise UmhsError(f"{what} failed with code {rc}: {msg}")


def ptr(t: Optional[torch.Tensor]):
    """Device address of a CUDA tensor (None -> NULL).  The tensor must be contiguous."""
    if t is None:
        return None
    if not t.is_cuda:
        raise UmhsError("libumhs_sm100 takes CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise UmhsError("non-contiguous tensor passed to a kernel")
    return t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()
