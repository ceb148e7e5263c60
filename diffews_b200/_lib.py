"""ctypes binding of the C ABI declared in include/diffews_b200.h.

The product path has no fallback: if the shared library is missing or does not export a symbol, importing this
module raises; if the device is not sm_100, every compute call returns DFW_ERR_ARCH and `check()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdiffews_b200.so")

DFW_OK, DFW_ERR_INVALID, DFW_ERR_CUDA, DFW_ERR_ARCH = 0, -1, -2, -3
EPI_OUT_F32, EPI_RES_F32, EPI_GEGLU, EPI_SILU, EPI_F16 = 1, 2, 4, 8, 16

# process-wide options (include/diffews_b200.h DFW_OPT_*)
(OPT_PDL, OPT_T128, OPT_T128_MAXC, OPT_HALO, OPT_GN_CTAS_PER_SM, OPT_PREPROC_TWO_PASS, OPT_ATTN_V2, OPT_SEG_HEAD,
 OPT_ATTN_BWD_UNFUSED, OPT_ATTN_V4, OPT_B_RESIDENT) = range(11)

_vp, _i, _f, _ll, _d = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_double

# name -> (restype, argtypes): mirrors include/diffews_b200.h one to one (tests/test_abi.py checks the header).
SIGNATURES = {
    "dfw_version": (_i, []),
    "dfw_device_ok": (_i, []),
    "dfw_launch_count": (_ll, []),
    "dfw_set_option": (_i, [_i, _i]),
    "dfw_get_option": (_i, [_i]),
    "dfw_conv2d_igemm": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "dfw_linear": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "dfw_upconv2x_igemm": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "dfw_gn_partial_floats": (_ll, [_i]),
    "dfw_conv_gnstats_supported": (_i, [_i, _i, _i, _i]),
    "dfw_conv2d_igemm_gnstats": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp]),
    "dfw_groupnorm_from_partial": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "dfw_gn_scale_shift": (_i, [_vp, _i, _vp, _vp, _vp, _i, _ll, _i, _i, _f, _vp]),
    "dfw_groupnorm_bwd_workspace_bytes": (_ll, [_i, _i, _i, _i]),
    "dfw_groupnorm_silu_bwd": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp, _vp]),
    "dfw_conv_gnin_supported": (_i, [_i, _i, _i, _i, _i, _i]),
    "dfw_conv_t128_eligible": (_i, [_i, _i, _i, _i, _i, _i]),
    "dfw_conv2d_igemm_gnin": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "dfw_conv_gnin_scratch_bytes": (_ll, []),
    "dfw_bmm_nt": (_i, [_vp, _vp, _ll, _ll, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp]),
    "dfw_attn_bwd_workspace_bytes": (_ll, [_i, _i, _i, _i, _i]),
    "dfw_attn_kvfused_bwd": (_i, [_vp, _ll, _i, _vp, _vp, _ll, _i, _vp, _vp, _ll, _i, _vp, _vp, _ll, _i, _vp, _vp, _vp, _vp,
                                  _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp, _vp]),
    "dfw_attn_kvfused_fwd_lse": (_i, [_vp, _ll, _i, _vp, _vp, _ll, _i, _vp, _vp, _ll, _i, _vp, _ll, _i, _i, _i, _i, _i,
                                      _i, _f, _i, _vp, _vp]),
    "dfw_attn_kvfused_fwd": (_i, [_vp, _ll, _i, _vp, _vp, _ll, _i, _vp, _vp, _ll, _i, _vp, _ll, _i, _i, _i, _i, _i,
                                  _i, _f, _i, _vp]),
    "dfw_cross_attn_fwd": (_i, [_vp, _vp, _vp, _ll, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    "dfw_cross_attn_collapsed": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _ll, _i, _i, _i, _vp]),
    "dfw_groupnorm_workspace_bytes": (_ll, [_i, _i, _i, _i]),
    "dfw_groupnorm_silu": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp, _vp]),
    "dfw_layernorm": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _f, _vp]),
    "dfw_layernorm_bwd_workspace_bytes": (_ll, [_i, _i]),
    "dfw_layernorm_bwd": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _f, _vp, _vp]),
    "dfw_softmax_rows": (_i, [_vp, _vp, _i, _i, _i, _f, _vp]),
    "dfw_split3_16": (_i, [_vp, _ll, _vp, _ll, _i, _i, _i, _vp]),
    "dfw_groupnorm_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    "dfw_layernorm_f32": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _f, _vp]),
    "dfw_softmax_rows_f32": (_i, [_vp, _vp, _i, _i, _f, _vp]),
    "dfw_geglu_f32": (_i, [_vp, _vp, _ll, _i, _vp]),
    "dfw_attn_f32": (_i, [_vp, _ll, _ll, _vp, _vp, _ll, _ll, _vp, _vp, _ll, _ll, _vp, _ll, _ll, _i, _i, _i, _i, _i, _f,
                          _vp]),
    "dfw_upsample2x_nhwc": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp]),
    "dfw_concat_channels": (_i, [_vp, _vp, _vp, _ll, _i, _i, _i, _vp]),
    "dfw_cast_f32_to_16": (_i, [_vp, _vp, _i, _ll, _vp]),
    "dfw_im2col3x3_small": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "dfw_conv3x3_small_cin": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "dfw_pointwise_small": (_i, [_vp, _ll, _ll, _ll, _vp, _vp, _f, _f, _vp, _ll, _ll, _ll, _i, _i, _i, _i, _vp]),
    "dfw_nhwc_f32_to_nchw_f32": (_i, [_vp, _i, _vp, _i, _i, _i, _f, _f, _f, _f, _vp]),
    "dfw_seg_post": (_i, [_vp, _i, _vp, _vp, _i, _i, _vp]),
    "dfw_seg_head_weight_u32": (_ll, []),
    "dfw_seg_head_prepare_weights": (_i, [_vp, _i, _vp]),
    "dfw_seg_head_u8": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "dfw_rthres_workspace_bytes": (_ll, [_i]),
    "dfw_rthres_iou_hist": (_i, [_vp, _i, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "dfw_iou_accumulate": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "dfw_preproc_workspace_bytes": (_ll, [_i, _i, _i, _i, _i]),
    "dfw_resize_normalize_u8": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _i, _f, _f, _vp, _ll, _vp]),
    "dfw_mask_nearest": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp]),
    "dfw_grad_norm_clip_coef": (_i, [_vp, _vp, _vp, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "dfw_adamw_step": (_i, [_vp, _vp, _vp, _i, _i, _d, _d, _d, _d, _d, _i, _vp, _i, _vp]),
    "dfw_geglu_bwd": (_i, [_vp, _vp, _vp, _i, _ll, _i, _vp]),
    "dfw_mse_workspace_floats": (_ll, []),
    "dfw_mse_loss": (_i, [_vp, _vp, _ll, _f, _vp, _vp, _vp, _vp]),
    "dfw_conv_wgrad_workspace_bytes": (_ll, [_i, _i, _i, _i, _i, _i, _i]),
    "dfw_conv_wgrad": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp, _vp]),
    "dfw_weight_permute": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "dfw_colsum_chunks": (_i, [_ll, _i]),
    "dfw_colsum": (_i, [_vp, _i, _vp, _ll, _i, _i, _f, _i, _vp, _vp]),
    "dfw_downsum2x_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "dfw_split_channels": (_i, [_vp, _vp, _vp, _ll, _i, _i, _vp]),
    "dfw_geglu_fwd": (_i, [_vp, _vp, _i, _ll, _i, _vp]),
    "dfw_nchw_f32_to_nhwc16_pad": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp]),
}


class DfwError(RuntimeError):
    pass


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU / PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:  # pragma: no cover
            raise ImportError(f"libdiffews_b200.so does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()

_ERR = {DFW_ERR_INVALID: "invalid argument", DFW_ERR_CUDA: "CUDA error", DFW_ERR_ARCH: "device is not sm_100 (B200)"}


def check(rc: int, what: str) -> None:
    if rc != DFW_OK:
        raise DfwError(f"{what} failed: {_ERR.get(rc, rc)} (see stderr)")
