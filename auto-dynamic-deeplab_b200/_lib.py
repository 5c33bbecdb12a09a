"""ctypes binding of libadd_b200.so (see include/add_b200.h).  There is NO fallback: if the
shared library is missing or a symbol is absent, importing this module raises."""
from __future__ import annotations

import ctypes
from ctypes import c_int, c_int32, c_int64, c_uint32, c_void_p, c_float, c_char_p, POINTER
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libadd_b200.so"

ADD_F32, ADD_BF16 = 0, 1
RELU_IN, RELU_OUT, ACCUMULATE = 1, 2, 4


class AddTensor(ctypes.Structure):
    """add_tensor_t — NHWC activation view (include/add_b200.h)."""
    _fields_ = [("ptr", c_void_p), ("n", c_int32), ("h", c_int32), ("w", c_int32), ("c", c_int32),
                ("pix_stride", c_int32), ("dtype", c_int32)]


class AddError(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  add_b200 has no CPU or PyTorch fallback.")
    return ctypes.CDLL(str(LIB_PATH))


lib = _load()

TP = POINTER(AddTensor)
_SIGS = {
    "add_status_string": (c_char_p, [c_int]),
    "add_last_cuda_error": (c_char_p, []),
    "add_version": (c_int, []),
    "add_device_sm_count": (c_int, []),
    "add_set_pdl": (c_int, [c_int]),
    "add_set_persistent_grid_pct": (c_int, [c_int]),
    "add_nchw_to_nhwc": (c_int, [c_void_p, c_int, TP, c_void_p]),
    "add_nhwc_to_nchw": (c_int, [TP, c_void_p, c_void_p]),
    "add_conv2d_fwd": (c_int, [TP, TP, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_uint32, c_void_p]),
    "add_conv2d_tc_packed_bytes": (c_int64, [c_int, c_int, c_int, c_int]),
    "add_conv2d_tc_pack": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "add_conv2d_tc_fwd": (c_int, [TP, TP, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_uint32, c_void_p]),
    "add_aspp_pool_bias_fwd": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "add_stem_tc_packed_bytes": (c_int64, []),
    "add_stem_tc_pack": (c_int, [c_void_p, c_void_p]),
    "add_stem_conv3x3s2_nchw_fwd": (c_int, [c_void_p, c_int, c_int, c_int, TP, c_void_p, c_void_p, c_uint32, c_void_p]),
    "add_conv2d_tc_set_halo_mode": (c_int, [c_int]),
    "add_sepconv_half_fwd": (c_int, [TP, TP, c_void_p, c_void_p, c_void_p, c_int, c_uint32, c_void_p]),
    "add_sepconv_half_tc_fwd": (c_int, [TP, TP, c_void_p, c_void_p, c_void_p, c_int, c_uint32, c_void_p]),
    "add_depthwise_fwd": (c_int, [TP, TP, c_void_p, c_int, c_uint32, c_void_p]),
    "add_sepconv_tc_set_mode": (c_int, [c_int]),
    "add_bilinear_set_mode": (c_int, [c_int]),
    "add_bn_stats_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int]),
    "add_bn_stats_fwd": (c_int, [TP, c_void_p, c_void_p, c_int64, c_void_p]),
    "add_bn_finalize": (c_int, [c_void_p, c_void_p, c_float, c_int, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p]),
    "add_bn_apply_fwd": (c_int, [TP, TP, c_void_p, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p]),
    "add_pool3x3_fwd": (c_int, [TP, TP, c_int, c_int, c_uint32, c_void_p]),
    "add_scale_fwd": (c_int, [TP, TP, c_float, c_int, c_uint32, c_void_p]),
    "add_bilinear_fwd": (c_int, [TP, TP, c_uint32, c_void_p]),
    "add_gather_images": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p]),
    "add_gather_images_view": (c_int, [TP, TP, c_void_p, c_void_p]),
    "add_global_avgpool_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int]),
    "add_global_avgpool_fwd": (c_int, [TP, c_void_p, c_uint32, c_void_p, c_int64, c_void_p]),
    "add_edm_mlp_fwd": (c_int, [c_void_p, c_int] + [c_void_p] * 6 + [c_void_p, c_void_p]),
    "add_upsample_logits_nchw": (c_int, [TP, c_void_p, c_int, c_int, c_void_p]),
    "add_head_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int]),
    "add_upsample_argmax_fwd": (c_int, [TP, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "add_upsample_argmax_u8_fwd": (c_int, [TP, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "add_normalize_u8_hwc_to_nchw": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int] + [ctypes.c_double] * 6 + [c_void_p]),
    "add_encode_pad_labels_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "add_normalize_pad_u8_hwc_to_nchw": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int] + [ctypes.c_double] * 6 + [c_void_p]),
    "add_decode_segmap": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p]),
    "add_relu_mask_bwd": (c_int, [TP, TP, c_void_p]),
    "add_conv2d_wgrad_workspace_bytes": (c_int64, [c_int] * 7),
    "add_conv2d_wgrad": (c_int, [TP, TP, c_void_p, c_int, c_int, c_int, c_int, c_int, c_uint32, c_void_p, c_int64, c_void_p]),
    "add_conv2d_dgrad": (c_int, [TP, c_void_p, TP, c_int, c_int, c_int, c_int, c_int, c_uint32, c_void_p]),
    "add_depthwise_wgrad_workspace_bytes": (c_int64, [c_int] * 5),
    "add_depthwise_wgrad": (c_int, [TP, TP, c_void_p, c_int, c_uint32, c_void_p, c_int64, c_void_p]),
    "add_bn_bwd_workspace_bytes": (c_int64, [c_int] * 4),
    "add_bn_bwd_reduce": (c_int, [TP, TP, c_void_p, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p, c_void_p, c_int64, c_void_p]),
    "add_bn_bwd_apply": (c_int, [TP, TP, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_double, c_void_p, c_void_p, c_uint32, TP, c_void_p]),
    "add_bilinear_bwd_tables": (c_int, [c_int, c_int] + [c_void_p] * 6 + [c_void_p]),
    "add_bilinear_bwd": (c_int, [TP, TP] + [c_void_p] * 12 + [c_uint32, c_void_p]),
    "add_pool3x3_bwd": (c_int, [TP, TP, TP, c_int, c_int, c_uint32, c_void_p]),
    "add_ce_loss_workspace_bytes": (c_int64, [c_int] * 3),
    "add_ce_loss_fwd_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int64, c_int64, c_int64, c_int, c_int64, c_void_p,
                                    c_float, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "add_sgd_nesterov": (c_int, [c_void_p, c_int, c_int64, c_float, c_void_p, c_float, c_float, c_int, c_int, c_void_p]),
    "add_peer_allreduce": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_uint32, c_void_p, c_int64, c_void_p, c_void_p]),
    "add_mixed_light_fwd": (c_int, [TP, TP, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p]),
    "add_weighted_sum_fwd": (c_int, [c_void_p, c_int, c_void_p, TP, c_void_p]),
    "add_weighted_sum_workspace_bytes": (c_int64, [c_int] * 4),
    "add_weighted_sum_bwd": (c_int, [TP, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "add_softmax_rows_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "add_softmax_rows_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "add_widen_labels_u8": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "add_confusion_workspace_bytes": (c_int64, [c_int64, c_int]),
    "add_confusion_matrix": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64, c_void_p]),
    "add_confusion_set_impl": (c_int, [c_int]),
    "add_confidence_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "add_confidence_nchw": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_int64, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)

for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)          # AttributeError here = the .so does not match the header
    _fn.restype = _res
    _fn.argtypes = _args


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib.add_status_string(int(status)).decode()
        if int(status) == -3:
            msg += " [" + lib.add_last_cuda_error().decode() + "]"
        raise AddError(f"libadd_b200 {what}: {msg} (status {status})")
