"""ctypes binding of libmobi_b200.so (the C ABI declared in include/mobi_b200.h).

There is deliberately no fallback: if the CUDA library is missing or fails to load, every op raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# MOBI_B200_LIB: an alternative build of the SAME library (kernel experiments compiled with other -D flags); never a fallback
LIB_PATH = os.environ.get("MOBI_B200_LIB") or os.path.join(_HERE, "libmobi_b200.so")

DT_BF16, DT_F32 = 0, 1
EPI_PLAIN, EPI_GEGLU, EPI_HEADS, EPI_HEADS_T, EPI_QKV, EPI_KV, EPI_GEGLU2 = 0, 1, 2, 3, 4, 5, 6
EPI_QKV_ROW, EPI_KV_ROW = 7, 8

_vp, _i64, _i32, _f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_float


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", _vp), ("B", _vp), ("out", _vp), ("out2", _vp), ("out3", _vp),
        ("bias", _vp), ("row_bias", _vp), ("residual", _vp),
        ("M", _i64), ("N", _i64), ("K", _i64),
        ("lda", _i64), ("ldb", _i64), ("ldo", _i64),
        ("rows_per_group", _i64), ("ld_row_bias", _i64),
        ("out_dtype", _i32), ("res_dtype", _i32), ("epilogue", _i32), ("act", _i32),
        ("heads", _i32), ("head_dim", _i32), ("tokens", _i32),
        ("out_seg", _i64), ("out_seg_stride", _i64), ("out_seg_offset", _i64),
        ("conv", _i32), ("n_img", _i32), ("H", _i32), ("W", _i32), ("C", _i32),
        ("KH", _i32), ("KW", _i32), ("pad_h", _i32), ("pad_w", _i32), ("tile_n", _i32), ("kernel", _i32),
        ("batch", _i32), ("a_batch_stride", _i64), ("b_batch_stride", _i64), ("out_batch_stride", _i64),
        ("pair", _i32), ("a_mn_major", _i32), ("b_mn_major", _i32), ("atomic_out", _i32),
        ("batch_inner", _i32), ("a_batch2_stride", _i64), ("b_batch2_stride", _i64), ("out_batch2_stride", _i64),
        ("colstats", _vp),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", _vp), ("k", _vp), ("vt", _vp), ("out", _vp),
        ("batch", _i32), ("heads", _i32), ("head_dim", _i32), ("tq", _i32), ("tk", _i32),
        ("ld_out", _i64), ("kernel", _i32), ("v_rowmajor", _i32), ("lse", _vp),
    ]


class GroupNormArgs(C.Structure):
    _fields_ = [
        ("x1", _vp), ("x2", _vp), ("gamma", _vp), ("beta", _vp), ("out", _vp), ("out_concat", _vp),
        ("partials", _vp),
        ("n_img", _i32), ("hw", _i32), ("c1", _i32), ("c2", _i32), ("groups", _i32),
        ("in_dtype", _i32), ("silu", _i32), ("eps", _f32), ("out_dtype", _i32), ("force_two_pass", _i32),
        ("colstats1", _vp), ("colstats2", _vp),
    ]


class LayerNormArgs(C.Structure):
    _fields_ = [
        ("x", _vp), ("gamma", _vp), ("beta", _vp), ("out", _vp), ("add_vec", _vp),
        ("rows", _i64), ("C", _i32),
        ("seg", _i64), ("seg_stride", _i64), ("seg_offset", _i64), ("add_rows_per_vec", _i64),
        ("eps", _f32),
    ]


class LnDualSpec(C.Structure):
    _fields_ = [("gamma", _vp * 2), ("beta", _vp * 2), ("out", _vp * 2), ("mode", _i32 * 2), ("pair", _i32)]


class LnAdapterArgs(C.Structure):
    _fields_ = [
        ("x", _vp), ("add_vec", _vp), ("gamma", _vp), ("beta", _vp), ("Ug", _vp), ("sb", _vp), ("Z", _vp), ("zb", _vp),
        ("batch", _i32), ("tokens", _i32), ("C", _i32), ("eps", _f32), ("next", LnDualSpec),
    ]


class Im2colArgs(C.Structure):
    _fields_ = [
        ("x", _vp), ("out", _vp), ("in_dtype", _i32),
        ("n", _i32), ("h", _i32), ("w", _i32), ("c", _i32), ("kh", _i32), ("kw", _i32),
        ("stride", _i32), ("pad_top", _i32), ("pad_left", _i32), ("ho", _i32), ("wo", _i32), ("kpad", _i32),
    ]


class CtxAttnArgs(C.Structure):
    _fields_ = [
        ("xn", _vp), ("U", _vp), ("Z", _vp), ("zb", _vp), ("x", _vp),
        ("batch", _i32), ("tokens", _i32), ("C", _i32), ("heads", _i32), ("keys", _i32),
    ]


class SamplerArgs(C.Structure):
    _fields_ = [
        ("eps", _vp), ("x", _vp), ("noise", _vp), ("old1", _vp), ("old2", _vp), ("old3", _vp),
        ("e_out", _vp), ("x_prev", _vp), ("pred_x0", _vp),
        ("n", _i64), ("cfg", _i32),
        ("scale", _f32), ("c0", _f32), ("c1", _f32), ("c2", _f32), ("c3", _f32),
        ("sqrt_one_minus_at", _f32), ("sqrt_at", _f32), ("sqrt_a_prev", _f32), ("dir_coef", _f32),
        ("sigma_temp", _f32),
    ]


class AssembleArgs(C.Structure):
    _fields_ = [
        ("x", _vp), ("inpaint_image", _vp), ("inpaint_mask", _vp),
        ("blend_mask", _vp), ("blend_x0", _vp), ("blend_noise", _vp), ("x_in", _vp),
        ("B", _i32), ("hw", _i32), ("cfg", _i32), ("blend_c", _i32), ("rest_c", _i32),
        ("sqrt_ac", _f32), ("sqrt_1mac", _f32),
    ]


class LayerNormBwdArgs(C.Structure):
    _fields_ = [
        ("x", _vp), ("gamma", _vp), ("dy", _vp), ("dx", _vp), ("dgamma", _vp), ("dbeta", _vp),
        ("rows", _i64), ("C", _i32),
        ("seg", _i64), ("seg_stride", _i64), ("seg_offset", _i64),
        ("eps", _f32), ("dy_dtype", _i32), ("accumulate", _i32),
    ]


class GroupNormBwdArgs(C.Structure):
    _fields_ = [
        ("x1", _vp), ("x2", _vp), ("gamma", _vp), ("beta", _vp), ("dy", _vp), ("dres", _vp), ("dx1", _vp), ("dx2", _vp),
        ("n_img", _i32), ("hw", _i32), ("c1", _i32), ("c2", _i32), ("groups", _i32), ("silu", _i32), ("dy_dtype", _i32),
        ("eps", _f32),
    ]


class AttnSoftmaxBwdArgs(C.Structure):
    _fields_ = [
        ("S", _vp), ("dP", _vp), ("dS", _vp), ("dSt", _vp), ("Pt", _vp), ("stats", _vp),
        ("batch", _i32), ("tq", _i32), ("tk", _i32), ("dscale", _f32),
    ]


class AttnBwdTilesArgs(C.Structure):
    _fields_ = [
        ("q", _vp), ("k", _vp), ("v", _vp), ("d_o", _vp), ("stats", _vp), ("dS", _vp), ("dSt", _vp), ("Pt", _vp),
        ("heads", _i32), ("tokens", _i32), ("head_dim", _i32), ("stats_only", _i32), ("ld_do", _i64), ("dscale", _f32),
        ("batch_rows", _i32),
    ]


class AttnBwdFlashArgs(C.Structure):
    _fields_ = [
        ("q", _vp), ("k", _vp), ("v", _vp), ("d_o", _vp), ("stats", _vp), ("dq", _vp), ("dk", _vp), ("dv", _vp),
        ("ld_do", _i64), ("ld_dq", _i64), ("ld_dk", _i64), ("ld_dv", _i64),
        ("heads", _i32), ("tokens", _i32), ("head_dim", _i32), ("batch_rows", _i32), ("dscale", _f32),
    ]


class CtxAttnQspaceArgs(C.Structure):
    _fields_ = [
        ("q", _vp), ("k", _vp), ("v", _vp), ("o", _vp), ("d_o", _vp), ("dq", _vp), ("dk", _vp), ("dv", _vp),
        ("batch", _i32), ("tokens", _i32), ("C", _i32), ("heads", _i32), ("keys", _i32), ("backward", _i32),
    ]


class LatentInputArgs(C.Structure):
    _fields_ = [
        ("moments_gt", _vp), ("noise_gt", _vp), ("moments_inpaint", _vp), ("noise_inpaint", _vp), ("mask", _vp), ("out", _vp),
        ("n", _i32), ("hs", _i32), ("ws", _i32), ("hm", _i32), ("wm", _i32), ("S", _i32), ("left", _i32), ("pad", _i32),
        ("row_stride", _i32), ("row_offset", _i32), ("scale", _f32),
    ]


class RangeMapArgs(C.Structure):
    _fields_ = [
        ("in_", _vp), ("out", _vp), ("min_d", _vp), ("max_d", _vp), ("n", _i64), ("in_stride", _i64), ("out_stride", _i64),
        ("batch", _i32), ("mode", _i32), ("clamp_input", _i32), ("alpha", C.c_double),
    ]


class RangeUndoArgs(C.Structure):
    _fields_ = [
        ("crop", _vp * 2), ("orig", _vp * 2), ("out", _vp * 2), ("min_d", _vp), ("max_d", _vp), ("crop_left", _vp),
        ("width_crop", _vp), ("crop_batch_stride", _i64), ("orig_batch_stride", _i64), ("out_batch_stride", _i64),
        ("batch", _i32), ("channels", _i32), ("crop_h", _i32), ("crop_w", _i32), ("H", _i32), ("W", _i32),
        ("zero_context", _i32), ("clamp_input", _i32), ("map", _i32 * 2), ("alpha", C.c_double),
    ]


class Range2PcdArgs(C.Structure):
    _fields_ = [
        ("depth", _vp), ("pitch", _vp), ("yaw", _vp), ("label", _vp), ("points", _vp), ("index", _vp), ("count", _vp),
        ("batch", _i32), ("H", _i32), ("W", _i32), ("point_stride", _i32), ("depth_min", _f32), ("depth_max", _f32),
    ]


class RangeCompositeArgs(C.Structure):
    _fields_ = [
        ("sample_depth", _vp), ("sample_int", _vp), ("depth_orig", _vp), ("int_orig", _vp), ("gt_mask", _vp), ("bbox", _vp),
        ("pitch", _vp), ("yaw", _vp), ("range_pred", _vp), ("pred_mask", _vp), ("points", _vp), ("count", _vp),
        ("batch", _i32), ("H", _i32), ("W", _i32), ("depth_min", _f32), ("depth_max", _f32),
    ]


# name -> (restype, argtypes); also the list of symbols include/mobi_b200.h declares
SIGNATURES = {
    "mobi_last_error": (C.c_char_p, []),
    "mobi_version": (C.c_int, []),
    "mobi_gemm": (C.c_int, [C.POINTER(GemmArgs), _vp]),
    "mobi_gemm_plan": (C.c_int, [C.POINTER(GemmArgs), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mobi_attention": (C.c_int, [C.POINTER(AttnArgs), _vp]),
    "mobi_groupnorm_scratch_bytes": (_i64, [_i32, _i32, _i32, _i32]),
    "mobi_groupnorm_launches": (_i32, [_i32, _i32, _i32, _i32, _i32]),
    "mobi_groupnorm": (C.c_int, [C.POINTER(GroupNormArgs), _vp]),
    "mobi_layernorm": (C.c_int, [C.POINTER(LayerNormArgs), _vp]),
    "mobi_ln_dual": (C.c_int, [_vp, C.POINTER(LnDualSpec), _i32, _i32, _i32, _f32, _vp]),
    "mobi_ln_adapter": (C.c_int, [C.POINTER(LnAdapterArgs), _vp]),
    "mobi_timestep_embedding": (C.c_int, [_vp, _vp, _i32, _i32, _f32, _vp]),
    "mobi_fourier_embed": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i64, _vp]),
    "mobi_silu": (C.c_int, [_vp, _i32, _vp, _i64, _vp]),
    "mobi_nchw_to_nhwc": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "mobi_nhwc_to_nchw": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _vp]),
    "mobi_upsample_nearest2x": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "mobi_im2col": (C.c_int, [C.POINTER(Im2colArgs), _vp]),
    "mobi_ctx_attention": (C.c_int, [C.POINTER(CtxAttnArgs), _vp]),
    "mobi_softmax_rows": (C.c_int, [_vp, _vp, _i64, _i32, _i64, _i64, _vp]),
    "mobi_sampler_update": (C.c_int, [C.POINTER(SamplerArgs), _vp]),
    "mobi_assemble_input": (C.c_int, [C.POINTER(AssembleArgs), _vp]),
    "mobi_add_f32": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "mobi_scale_f32": (C.c_int, [_vp, _f32, _vp, _i64, _vp]),
    "mobi_cast_bf16": (C.c_int, [_vp, _vp, _i64, _vp]),
    "mobi_cast_bf16_segments": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i32, _vp]),
    "mobi_scale_segments": (C.c_int, [_vp, _i64, _vp, _vp, _i32, _vp]),
    # training step
    "mobi_transpose_bf16": (C.c_int, [_vp, _i32, _vp, _i64, _i32, _i32, _i64, _i64, _i64, _i64, _vp]),
    "mobi_layernorm_bwd": (C.c_int, [C.POINTER(LayerNormBwdArgs), _vp]),
    "mobi_groupnorm_bwd": (C.c_int, [C.POINTER(GroupNormBwdArgs), _vp]),
    "mobi_geglu": (C.c_int, [_vp, _vp, _i64, _i64, _vp]),
    "mobi_geglu_bwd": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp]),
    "mobi_attn_softmax_bwd": (C.c_int, [C.POINTER(AttnSoftmaxBwdArgs), _vp]),
    "mobi_attn_softmax_bwd_lse": (C.c_int, [C.POINTER(AttnSoftmaxBwdArgs), _vp, _vp, _vp, _i64, _i32, _vp]),
    "mobi_attn_bwd_tiles": (C.c_int, [C.POINTER(AttnBwdTilesArgs), _vp]),
    "mobi_ctx_attn_qspace": (C.c_int, [C.POINTER(CtxAttnQspaceArgs), _vp]),
    "mobi_colsum": (C.c_int, [_vp, _i32, _i64, _i32, _i64, _i64, _vp, _vp]),
    "mobi_wgrad_small": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i64, _i64, _i64, _vp]),
    "mobi_scatter_add_rows": (C.c_int, [_vp, _i32, _vp, _i64, _i32, _i64, _i64, _i64, _vp]),
    "mobi_zero_insert2x": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _vp]),
    "mobi_sum2x2": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "mobi_q_sample": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "mobi_mse_grad": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f32, _vp]),
    "mobi_assemble_latent_input": (C.c_int, [C.POINTER(LatentInputArgs), _vp]),
    "mobi_bbox_renorm": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _i32, _vp]),
    "mobi_silu_bwd": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _i64, _vp]),
    "mobi_adamw": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _f32, _f32, _f32, _vp]),
    "mobi_adamw_segments": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _f32, _f32, _f32, _f32, _f32, _f32,
                                      _vp]),
    "mobi_attn_bwd_flash": (C.c_int, [C.POINTER(AttnBwdFlashArgs), _vp]),
    "mobi_range_map": (C.c_int, [C.POINTER(RangeMapArgs), _vp]),
    "mobi_range_undo_transforms": (C.c_int, [C.POINTER(RangeUndoArgs), _vp]),
    "mobi_range2pcd": (C.c_int, [C.POINTER(Range2PcdArgs), _vp]),
    "mobi_range_composite": (C.c_int, [C.POINTER(RangeCompositeArgs), _vp]),
    "mobi_points_in_boxes": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _i32, _vp]),
}

_lib = None


def load():
    """Loads the library (building is a separate, explicit step: `python -m mobi_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "mobi_b200: %s is missing. Build it with `python -m mobi_b200.build` (needs nvcc); "
            "there is no CPU or PyTorch fallback for this path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the .so is stale
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().mobi_last_error().decode("utf-8", "replace")
        raise RuntimeError("mobi_b200.%s failed (status %d): %s" % (what, status, msg))


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def dt(t):
    if t.dtype == torch.bfloat16:
        return DT_BF16
    if t.dtype == torch.float32:
        return DT_F32
    raise TypeError("mobi_b200: unsupported dtype %s" % t.dtype)
