"""ctypes binding of libfpg_b200.so (C ABI declared in include/fpg.h).

The library is the product's only compute path: there is no CPU fallback. Importing this module only loads the
shared object (that works without a GPU, so symbol/plan tests can run on CPU); every compute call needs a B200.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libfpg_b200.so")

FPG_MAX_TAPS = 64
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_TANH = 0, 1, 2, 3
DT_BF16, DT_FP32, DT_FP16 = 0, 1, 2  # fpg_act.fp32 / fpg_out_view.fp32 element types (FPG_DT_*)
ABI_VERSION = 2


class Tap(C.Structure):
    _fields_ = [("c0", C.c_int32), ("dx", C.c_int32), ("plane", C.c_int32), ("dy", C.c_int32)]


class TMap(C.Structure):
    _fields_ = [("base", C.c_void_p), ("rank", C.c_int32), ("swizzle_bytes", C.c_int32),
                ("dims", C.c_uint64 * 5), ("strides", C.c_uint64 * 4), ("box", C.c_uint32 * 5)]


class OutView(C.Structure):
    _fields_ = [("base", C.c_void_p), ("stride_n", C.c_int64), ("stride_y", C.c_int64), ("stride_x", C.c_int64),
                ("mul_y", C.c_int32), ("off_y", C.c_int32), ("mul_x", C.c_int32), ("off_x", C.c_int32),
                ("valid_h", C.c_int32), ("valid_w", C.c_int32), ("fp32", C.c_int32)]


class FpropDesc(C.Structure):
    _fields_ = [("a", TMap), ("b", TMap), ("cblk", C.c_int32), ("c_per_tap", C.c_int32), ("num_taps", C.c_int32),
                ("num_sub", C.c_int32), ("block_n", C.c_int32), ("n_blocks", C.c_int32), ("n_img", C.c_int32),
                ("tiles_y", C.c_int32), ("tiles_x", C.c_int32), ("tile_h", C.c_int32), ("tile_w", C.c_int32),
                ("act", C.c_int32), ("stages", C.c_int32), ("cta_pair", C.c_int32), ("bias", C.c_void_p), ("out", OutView),
                ("taps", Tap * FPG_MAX_TAPS), ("a1", TMap), ("tiles_y1", C.c_int32), ("tiles_x1", C.c_int32),
                ("tile_h1", C.c_int32), ("tile_w1", C.c_int32), ("x_org1", C.c_int32), ("stat_partial", C.c_void_p),
                ("stat_rows_per_img", C.c_int32), ("stat_row0", C.c_int32), ("inbwd_mode", C.c_int32),
                ("inbwd_has_add", C.c_int32), ("inbwd_h", C.c_int32), ("inbwd_w", C.c_int32),
                ("inbwd_halo", C.c_int32), ("inbwd_add_halo", C.c_int32), ("inbwd_z", TMap), ("inbwd_z1", TMap),
                ("inbwd_prev", TMap), ("inbwd_prev1", TMap), ("inbwd_add", TMap), ("inbwd_add1", TMap)]


class RowsDesc(C.Structure):
    _fields_ = [("a", TMap), ("b", TMap), ("cblk", C.c_int32), ("block_n", C.c_int32), ("rows", C.c_int32),
                ("cols", C.c_int32), ("dy0", C.c_int32), ("dx0", C.c_int32), ("tile_rows", C.c_int32),
                ("n_img", C.c_int32), ("tiles_y", C.c_int32), ("tiles_x", C.c_int32), ("act", C.c_int32),
                ("a_stages", C.c_int32), ("b_stages", C.c_int32), ("bias", C.c_void_p), ("out", OutView),
                ("tap_of", C.c_int16 * FPG_MAX_TAPS)]


class PackJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("rows", C.c_int32), ("taps", C.c_int32),
                ("cols", C.c_int32), ("rows_valid", C.c_int32), ("cols_valid", C.c_int32), ("dst_fp32", C.c_int32),
                ("src_stride_row", C.c_int64), ("src_stride_col", C.c_int64), ("src_tap", C.c_int8 * FPG_MAX_TAPS)]


class WgradDesc(C.Structure):
    _fields_ = [("x", TMap), ("y", TMap), ("x_ca", C.c_int32), ("y_ca", C.c_int32), ("x_atoms", C.c_int32),
                ("y_atoms", C.c_int32), ("x_groups", C.c_int32), ("y_groups", C.c_int32),
                ("x_taps_mode", C.c_int32), ("y_taps_mode", C.c_int32), ("x_ntaps", C.c_int32),
                ("y_ntaps", C.c_int32), ("n_img", C.c_int32), ("kt_y", C.c_int32), ("kt_x", C.c_int32),
                ("tile_h", C.c_int32), ("tile_w", C.c_int32), ("splits", C.c_int32), ("stages", C.c_int32),
                ("x_is_dy", C.c_int32), ("tap_on_x", C.c_int32), ("taps_r", C.c_int32), ("taps_s", C.c_int32),
                ("x_shift_atoms", C.c_int32), ("y_shift_atoms", C.c_int32), ("y_shifts", C.c_int32), ("y_sets", C.c_int32),
                ("cta_pair", C.c_int32), ("last_splits", C.c_int32), ("ws", C.c_void_p),
                ("x_taps", Tap * FPG_MAX_TAPS), ("y_taps", Tap * FPG_MAX_TAPS),
                ("x_tap_rs", C.c_int16 * FPG_MAX_TAPS), ("y_tap_rs", C.c_int16 * FPG_MAX_TAPS)]


class AdamPackLayer(C.Structure):
    _fields_ = [("p_off", C.c_int64), ("k", C.c_int32), ("c", C.c_int32), ("rs", C.c_int32), ("tk", C.c_int32),
                ("tc", C.c_int32), ("n_jobs", C.c_int32), ("job", C.c_int32 * 8)]


class AdamPackChunk(C.Structure):
    _fields_ = [("off", C.c_int64), ("copy_dst", C.c_void_p), ("count", C.c_int32), ("pad_", C.c_int32)]


class Act(C.Structure):
    """fpg_act: NHWC activation buffer descriptor."""
    _fields_ = [("data", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("c_stride", C.c_int32), ("halo", C.c_int32), ("fp32", C.c_int32)]


class ConvGeom(C.Structure):
    _fields_ = [("r", C.c_int32), ("s", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
                ("c_in", C.c_int32), ("c_out", C.c_int32)]


_P = C.POINTER
_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); kept in one table so the "exports every declared symbol" test can iterate it
SIGNATURES = {
    "fpg_abi_version": (C.c_int, []),
    "fpg_last_error": (C.c_char_p, []),
    "fpg_sm_count": (C.c_int, []),
    "fpg_igemm_fprop_launch": (C.c_int, [_P(FpropDesc), _vp]),
    "fpg_igemm_wgrad_launch": (C.c_int, [_P(WgradDesc), _vp]),
    "fpg_igemm_rows_launch": (C.c_int, [_P(RowsDesc), _vp]),
    "fpg_conv2d_rows_plan": (C.c_int, [_P(Act), _vp, _vp, C.c_int, _P(ConvGeom), _P(Act), C.c_int, C.c_int,
                                       _P(RowsDesc)]),
    "fpg_conv2d_fprop": (C.c_int, [_P(Act), _vp, _vp, C.c_int, _P(ConvGeom), _P(Act), _vp]),
    "fpg_conv2d_fprop_plan": (C.c_int, [_P(Act), _vp, _vp, C.c_int, _P(ConvGeom), _P(Act), C.c_int, _P(FpropDesc)]),
    "fpg_conv2d_dgrad": (C.c_int, [_P(Act), _vp, _vp, C.c_int, _P(ConvGeom), _P(Act), _vp]),
    "fpg_conv2d_dgrad_plan": (C.c_int, [_P(Act), _vp, _vp, C.c_int, _P(ConvGeom), _P(Act), C.c_int, _P(FpropDesc),
                                        _P(C.c_int)]),
    "fpg_conv_stats_rows": (_i32, [_P(Act), _P(ConvGeom), _P(Act), C.c_int]),
    "fpg_igemm_s2cls_launch": (C.c_int, [_P(FpropDesc), _i32, _vp]),
    "fpg_conv2d_dgrad_launches": (_i32, [_P(Act), _P(ConvGeom), _P(Act)]),
    "fpg_conv2d_fprop_stats": (C.c_int, [_P(Act), _vp, _vp, C.c_int, _P(ConvGeom), _P(Act), _vp, _vp]),
    "fpg_conv2d_dgrad_stats": (C.c_int, [_P(Act), _vp, _vp, C.c_int, _P(ConvGeom), _P(Act), _vp, _vp]),
    "fpg_instnorm_stats_finalize": (C.c_int, [_vp, _i32, _i32, _i32, _i64, _f32, _vp, _vp]),
    "fpg_conv2d_wgrad": (C.c_int, [_P(Act), _P(Act), _P(ConvGeom), _vp, _i64, _i64, _i32, _i32, _vp, _vp]),
    "fpg_conv2d_wgrad_plan": (C.c_int, [_P(Act), _P(Act), _P(ConvGeom), C.c_int, _P(WgradDesc)]),
    "fpg_conv2d_wgrad_ws_bytes": (_i64, [_P(Act), _P(Act), _P(ConvGeom), C.c_int]),
    "fpg_pack_weights": (C.c_int, [_vp, _i64, _i64, _i32, _i32, _P(ConvGeom), _vp, _vp]),
    "fpg_pack_weights_dgrad": (C.c_int, [_vp, _i64, _i64, _i32, _i32, _P(ConvGeom), _vp, _vp]),
    "fpg_pack_jobs": (C.c_int, [_vp, _i64, _i64, _i32, _i32, _P(ConvGeom), _vp, _vp, _P(PackJob), _P(_i32)]),
    "fpg_pack_job_copy_f32": (C.c_int, [_vp, _i32, _vp, _i32, _P(PackJob)]),
    "fpg_pack_job_blocks": (_i32, [_P(PackJob)]),
    "fpg_pack_weights_batched": (C.c_int, [_vp, _vp, _vp, _i32, _vp]),
    "fpg_dgrad_class_info": (C.c_int, [_P(ConvGeom), C.c_int, _P(_i32), _P(_i32), _P(_i64), _P(_i32)]),
    "fpg_packed_weight_bytes": (_i64, [_P(ConvGeom)]),
    "fpg_packed_weight_dgrad_bytes": (_i64, [_P(ConvGeom)]),
    "fpg_bias_grad": (C.c_int, [_P(Act), _vp, _i32, _vp, _vp]),
    "fpg_instnorm_scratch_floats": (_i64, [_P(Act)]),
    "fpg_instnorm_stats": (C.c_int, [_P(Act), _f32, _vp, _vp, _vp, _vp]),
    "fpg_instnorm_apply": (C.c_int, [_P(Act), _vp, C.c_int, _P(Act), _P(Act), _P(Act), _vp]),
    "fpg_conv2d_dgrad_inbwd": (C.c_int, [_P(Act), _vp, _P(ConvGeom), _P(Act), _P(Act), _P(Act), _P(Act), _vp,
                                         _P(_i32), _vp]),
    "fpg_instnorm_bwd_sums_finalize": (C.c_int, [_vp, _i32, _i32, _i32, _i64, _vp, _vp]),
    "fpg_instnorm_bwd_apply": (C.c_int, [_P(Act), _P(Act), _vp, _vp, C.c_int, _P(Act), _vp]),
    "fpg_instnorm_bwd": (C.c_int, [_P(Act), _P(Act), _P(Act), _vp, C.c_int, _P(Act), _P(Act), _vp, _vp, _vp]),
    "fpg_batchnorm_apply": (C.c_int, [_P(Act), _vp, _vp, _vp, _vp, _f32, C.c_int, _P(Act), C.c_int, _P(Act), _vp]),
    "fpg_batchnorm_bwd": (C.c_int, [_P(Act), C.c_int, _P(Act), C.c_int, _vp, _f32, _P(Act), _vp, _vp, _vp, _P(Act),
                                    _vp, _vp, C.c_int, _vp, _vp]),
    "fpg_batchnorm_scratch_floats": (_i64, [_P(Act)]),
    "fpg_batchnorm_running_update": (C.c_int, [_vp, _i32, _i64, _f32, _f32, _vp, _vp, _vp]),
    "fpg_maxpool2": (C.c_int, [_P(Act), _P(Act), _vp]),
    "fpg_dropout_mask": (C.c_int, [_vp, _i64, C.c_uint64, _f32, _vp]),
    "fpg_dropout_mask_dev": (C.c_int, [_vp, _i64, _vp, C.c_uint64, _f32, _vp]),
    "fpg_ssim_scratch_bytes": (_i64, [_i32, _i32, _i32, _i32]),
    "fpg_ssim_stats": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _f32, _vp, _vp, _vp]),
    "fpg_avgpool2_f32": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp]),
    "fpg_sq_err_scratch_bytes": (_i64, []),
    "fpg_sq_err_sum": (C.c_int, [_vp, _vp, _i64, _f32, _f32, _vp, _vp, _vp]),
    "fpg_resize_aa_scratch_bytes": (_i64, [_i32, _i32, _i32, _i32, _i32]),
    "fpg_resize_bicubic_aa": (C.c_int, [_vp, _i32, _i32, _i32, _P(_i32), _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "fpg_tile_gather": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _i32, _i32, _f32, _f32, _vp, _vp]),
    "fpg_act_bwd": (C.c_int, [_P(Act), _P(Act), C.c_int, _P(Act), _vp]),
    "fpg_halo_fold": (C.c_int, [_P(Act), _P(Act), _P(Act), _vp]),
    "fpg_blend_fwd": (C.c_int, [_P(Act), _P(Act), _P(Act), _i32, _P(Act), _i32, _vp, _vp, _vp]),
    "fpg_norm_split_f32": (C.c_int, [_P(Act), C.c_int, _f32, C.c_int, _vp, _vp, _P(Act), _vp]),
    "fpg_pack_nchw_split": (C.c_int, [_vp, _i32, _P(Act), _vp]),
    "fpg_blend_bwd": (C.c_int, [_vp, _P(Act), _i32, _P(Act), _P(Act), _P(Act), _P(Act), _P(Act), _vp, _vp]),
    "fpg_mse_const_loss": (C.c_int, [_P(Act), _f32, _f32, _f32, _vp, _P(Act), _vp, _vp, _vp]),
    "fpg_l1_loss": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _f32, _f32, _vp, _vp, C.c_int, _vp, _vp]),
    "fpg_pack_nchw": (C.c_int, [_vp, _i32, _i32, _P(Act), _i32, C.c_int, _vp]),
    "fpg_pack_paired_inputs": (C.c_int, [_vp, _i32, _vp, _i32, _P(Act), _P(Act), _P(Act), _vp]),
    "fpg_space_to_depth16": (C.c_int, [_P(Act), _P(Act), _vp]),
    "fpg_add_f32": (C.c_int, [_vp, _vp, _i64, _vp]),
    "fpg_history_exchange": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "fpg_unpack_nchw": (C.c_int, [_P(Act), _i32, _vp, _i32, C.c_int, _vp]),
    "fpg_tanh_bwd_pack": (C.c_int, [_vp, _P(Act), _i32, _P(Act), _vp]),
    "fpg_adam_step": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _i32, _f32, _vp]),
    "fpg_adam_step_dev": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _vp, _f32, _vp]),
    "fpg_adam_pack_tile": (C.c_int, [_i32, _P(_i32), _P(_i32)]),
    "fpg_adam_pack_step": (C.c_int, [_vp, _P(_vp), _i32, _vp, _vp, _f32, _f32, _f32, _vp, _f32, _vp, _vp, _vp, _vp, _vp,
                                     _vp, _i32, _vp]),
    "fpg_adam_prepare_dev": (C.c_int, [_vp, _f32, _f32, _vp]),
    "fpg_peer_alloc": (C.c_int, [_P(_vp), _i64]),
    "fpg_peer_free": (C.c_int, [_vp]),
    "fpg_peer_export": (C.c_int, [_vp, _vp]),
    "fpg_peer_open": (C.c_int, [_vp, _P(_vp)]),
    "fpg_peer_close": (C.c_int, [_vp]),
    "fpg_peer_copy": (C.c_int, [_vp, _vp, _i64, _vp]),
    "fpg_peer_signal": (C.c_int, [_P(_vp), _i32, _vp, C.c_uint32, C.c_uint32, _vp]),
    "fpg_peer_wait": (C.c_int, [_vp, _i32, _i32, _vp, C.c_uint32, _vp, _f32, _vp]),
    "fpg_peer_push": (C.c_int, [_vp, _P(_vp), _i32, _i64, _vp]),
    "fpg_adam_step_dev_multi": (C.c_int, [_vp, _P(_vp), _i32, _vp, _vp, _i64, _f32, _f32, _f32, _vp, _f32, _vp, _vp]),
    "fpg_flood_mask": (C.c_int, [_vp, _vp, _i64, _vp]),
    "fpg_confusion_counts": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
}

_lib = None


class FpgError(RuntimeError):
    pass


def load():
    """Load libfpg_b200.so (raises if it has not been built: there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FpgError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C flood-prediction-gan_b200/csrc`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.fpg_abi_version() != ABI_VERSION:
        raise FpgError("libfpg_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().fpg_last_error().decode("utf-8", "replace")
        raise FpgError(f"{what} failed with code {rc}: {msg}")


def call(name, *args):
    """Call a status-returning entry point and raise FpgError on a non-zero status."""
    rc = getattr(load(), name)(*args)
    check(rc, name)
