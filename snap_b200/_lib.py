"""ctypes binding of libsnapb200.so (the C ABI declared in include/snapb200.h).

There is deliberately no fallback: if the shared library is missing, or a compute entry point is
called without a CUDA device, the call raises.
"""
from __future__ import annotations

import ctypes as C
import pathlib

_LIB_PATH = pathlib.Path(__file__).resolve().parent / "libsnapb200.so"
_lib = None


class SnapB200Error(RuntimeError):
    pass


class GemmParams(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_rows", C.c_longlong), ("a_cols", C.c_int), ("a_ld", C.c_longlong),
        ("b", C.c_void_p), ("b_rows", C.c_longlong), ("b_cols", C.c_int), ("b_ld", C.c_longlong),
        ("m_rows", C.c_longlong), ("n", C.c_int), ("num_seg", C.c_int), ("seg_k", C.c_int),
        ("a_col0", C.c_int), ("seg_off", C.c_int * 9),
        ("out", C.c_void_p), ("ldo", C.c_longlong), ("out_f32", C.c_int),
        ("residual", C.c_void_p), ("ldr", C.c_longlong),
        ("bias", C.c_void_p), ("row_mask", C.c_void_p), ("relu", C.c_int),
        ("remap", C.c_int), ("rm_R", C.c_int), ("rm_C", C.c_int), ("rm_r0", C.c_int),
        ("rm_c0", C.c_int), ("rm_Ho", C.c_int), ("rm_Wo", C.c_int),
        ("bn", C.c_int),
        ("gn_acc", C.c_void_p), ("gn_acc_relu", C.c_void_p), ("gn_rows_per_img", C.c_longlong),
        ("gn_replica_stride", C.c_int),
    ]


def lib() -> C.CDLL:
    """Load the library once. Raises if it has not been built (`python -m snap_b200.build`)."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise SnapB200Error(
                f"{_LIB_PATH} is missing: build it with `python -m snap_b200.build` "
                "(there is no CPU/PyTorch fallback for the hot path)")
        l = C.CDLL(str(_LIB_PATH))
        l.snapb200_last_error.restype = C.c_char_p
        l.snapb200_launch_count.restype = C.c_longlong
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise SnapB200Error(f"libsnapb200 error {rc}: {lib().snapb200_last_error().decode()}")


def launch_count() -> int:
    return int(lib().snapb200_launch_count())


def launch_count_reset() -> None:
    lib().snapb200_launch_count_reset()


# Every symbol include/snapb200.h declares; tests/test_abi.py checks the .so exports each of them.
EXPORTED_SYMBOLS = [
    "snapb200_last_error", "snapb200_version", "snapb200_launch_count",
    "snapb200_launch_count_reset", "snapb200_gemm_bf16", "snapb200_conv_gn_bf16", "snapb200_selftest_shifted_desc",
    "snapb200_conv3x3_halo_supported", "snapb200_conv3x3_halo_bf16",
]


class Conv3x3Params(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("n_img", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int),
        ("b", C.c_void_p), ("b_ld", C.c_longlong), ("n", C.c_int), ("out", C.c_void_p), ("ldo", C.c_longlong),
        ("gn_acc", C.c_void_p), ("gn_replica_stride", C.c_int),
    ]


class ConvGnParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("n_img", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int),
        ("acc", C.c_void_p), ("replica_stride", C.c_int), ("scale", C.c_void_p), ("bias", C.c_void_p),
        ("pre_relu", C.c_int), ("post_relu", C.c_int), ("taps", C.c_int), ("stride", C.c_int),
        ("b", C.c_void_p), ("b_rows", C.c_longlong), ("b_cols", C.c_int), ("b_ld", C.c_longlong),
        ("n", C.c_int), ("out", C.c_void_p), ("ldo", C.c_longlong), ("residual", C.c_void_p),
        ("ldr", C.c_longlong), ("gn_acc", C.c_void_p), ("gn_acc_relu", C.c_void_p),
        ("gn_replica_stride", C.c_int),
    ]


class RootConvParams(C.Structure):
    _fields_ = [
        ("packed", C.c_void_p), ("n_img", C.c_int), ("Hq", C.c_int), ("Wq", C.c_int), ("cp", C.c_int),
        ("KH", C.c_int), ("stride", C.c_int), ("Ho", C.c_int), ("Wo", C.c_int),
        ("b", C.c_void_p), ("n", C.c_int), ("out", C.c_void_p), ("ldo", C.c_longlong),
        ("gn_acc", C.c_void_p), ("gn_replica_stride", C.c_int),
    ]


class WeightDesc(C.Structure):
    _fields_ = [("w", C.c_void_p), ("out", C.c_void_p), ("partial", C.c_void_p),
                ("K", C.c_int), ("Cout", C.c_int), ("ldb", C.c_int), ("standardize", C.c_int),
                ("ksplit", C.c_int), ("pad_", C.c_int)]


MAX_VIEWS = 8            # all-views kernels (V <= top_k_view_selection) and top_k itself
MAX_SELECT_VIEWS = 32    # view-selection kernel (V > top_k)


class LiftView(C.Structure):
    _fields_ = [("Rinv", C.c_float * 9), ("tinv", C.c_float * 3), ("f", C.c_float * 2),
                ("c", C.c_float * 2), ("wh", C.c_float * 2), ("k_radial", C.c_float * 3),
                ("tan_half_fov", C.c_float), ("fisheye", C.c_int)]


class LiftParams(C.Structure):
    _fields_ = [("V", C.c_int), ("Hf", C.c_int), ("Wf", C.c_int), ("CF", C.c_int), ("D", C.c_int),
                ("S", C.c_int), ("X", C.c_int), ("Y", C.c_int), ("Z", C.c_int),
                ("depth_min", C.c_float), ("depth_max", C.c_float), ("inv_log_range", C.c_float),
                ("stats_ld", C.c_int), ("xy_paired", C.c_int),
                ("no_variance", C.c_int), ("add_minmax", C.c_int)]


EXPORTED_SYMBOLS += [
    "snapb200_std_weights_batched", "snapb200_root_im2col", "snapb200_root_pack_image", "snapb200_root_pack_weights",
    "snapb200_root_conv_bf16", "snapb200_maxpool3x3s2",
    "snapb200_gn_stats", "snapb200_gn_apply", "snapb200_upsample2x",
    "snapb200_crop_relu", "snapb200_lift_gather_pool", "snapb200_lift_select_pool", "snapb200_lift_fused", "snapb200_lift_fused_scratch_bytes", "snapb200_lift_fused_batched", "snapb200_lift_fused_batched_scratch_bytes", "snapb200_lift_observe", "snapb200_lift_pool_observations", "snapb200_vertical_max", "snapb200_vertical_pool", "snapb200_mask_rows", "snapb200_confidence", "snapb200_valid_any", "snapb200_match_head", "snapb200_match_head_ex",
    "snapb200_fuse_max", "snapb200_xcorr_padded_cols", "snapb200_xcorr_padded_rotations", "snapb200_rot_templates",
    "snapb200_xcorr_pad_map", "snapb200_xcorr_count", "snapb200_xcorr_scores", "snapb200_xcorr_scores_sw", "snapb200_xcorr_scores_rows", "snapb200_xcorr_scores_rows_workspace",
    "snapb200_loc_softmax_stats", "snapb200_loc_point_weights", "snapb200_loc_sample", "snapb200_loc_ransac_poses",
    "snapb200_loc_refine_poses", "snapb200_loc_pose_scoring_workspace", "snapb200_loc_pose_scoring",
    "snapb200_argmax_rows", "snapb200_loc_nll", "snapb200_sem_loss",
    "snapb200_sem_loss_grad", "snapb200_relu_bwd", "snapb200_dense_wgrad_workspace", "snapb200_dense_wgrad",
    "snapb200_adam_step", "snapb200_cast_pad_bf16", "snapb200_sem_labels",
    "snapb200_gn_backward", "snapb200_upsample2x_backward", "snapb200_maxpool3x3s2_backward", "snapb200_wt_segments", "snapb200_stdconv_backward",
    "snapb200_lift_gather_pool_backward", "snapb200_lift_select_pool_backward", "snapb200_vertical_max_backward",
    "snapb200_match_head_backward", "snapb200_fuse_max_backward",
    "snapb200_loc_nll_backward", "snapb200_loc_pose_scoring_backward",
]


SEM_OUT = 40


class SemLossParams(C.Structure):
    _fields_ = [("B", C.c_int), ("cells", C.c_int), ("num_area", C.c_int), ("num_excl", C.c_int),
                ("num_indep", C.c_int), ("ld", C.c_int)]


class LocScoreParams(C.Structure):
    _fields_ = [("B", C.c_int), ("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("P", C.c_int),
                ("cell_size", C.c_float), ("mask_out_of_bounds", C.c_int), ("i_xy_batched", C.c_int)]
