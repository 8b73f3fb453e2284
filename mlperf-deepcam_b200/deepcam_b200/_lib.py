"""ctypes binding of include/deepcam_b200.h.

The product path has no CPU or library fallback: if the shared library cannot be loaded (and cannot be
built), importing the kernels raises.  ctypes releases the GIL during every call.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p)

from . import build as _build

DC_F32, DC_BF16 = 0, 1
DC_MAX_TAPS = 9
DC_PACK_TKN, DC_PACK_NTK, DC_PACK_NTK_CONVT2 = 0, 1, 2
DC_ABI_VERSION = 2
DC_CONV_WEIGHTS_STABLE, DC_CONV_HALO_PACK = 1, 2
DC_BN_RELU, DC_BN_TRAIN, DC_BN_IDENTITY, DC_BN_RES_WRITE, DC_BN_SUMS_READY, DC_BN_MASK_FROM_Y = 1, 2, 4, 8, 16, 32


class dc_view(Structure):
    _fields_ = [("ptr", c_void_p), ("n", c_int32), ("h", c_int32), ("w", c_int32), ("c", c_int32),
                ("sn", c_int64), ("sh", c_int64), ("sw", c_int64), ("sc", c_int64),
                ("dtype", c_int32), ("reserved", c_int32)]


class dc_conv_desc(Structure):
    _fields_ = [("ntaps", c_int32), ("dh", c_int32 * DC_MAX_TAPS), ("dw", c_int32 * DC_MAX_TAPS),
                ("wt", c_int32 * DC_MAX_TAPS), ("stride_h", c_int32), ("stride_w", c_int32),
                ("accumulate", c_int32), ("wtaps", c_int32),
                ("out_csplit", c_int32), ("flags", c_int32), ("out_split_off", c_int64)]


class dc_bn_params(Structure):
    _fields_ = [("gamma", c_void_p), ("beta", c_void_p), ("running_mean", c_void_p), ("running_var", c_void_p),
                ("sums", c_void_p), ("count", c_double), ("momentum", c_float), ("eps", c_float),
                ("flags", c_int32), ("reserved", c_int32)]


class dc_pack_job(Structure):
    _fields_ = [("src", c_void_p), ("dst", c_void_p), ("K", c_int32), ("N", c_int32), ("taps", c_int32),
                ("src_k_first", c_int32), ("layout", c_int32), ("K_pad", c_int32), ("N_pad", c_int32),
                ("dst_dtype", c_int32), ("block_start", c_int32), ("n_blocks", c_int32)]


class dc_adam_job(Structure):
    _fields_ = [("p", c_void_p), ("g", c_void_p), ("m", c_void_p), ("v", c_void_p), ("numel", c_int64),
                ("block_start", c_int32), ("n_blocks", c_int32)]


# name -> (restype, argtypes); must list every symbol declared in include/deepcam_b200.h
SIGNATURES = {
    "dc_abi_version": (c_int, []),
    "dc_last_error_string": (c_char_p, []),
    "dc_device_supports_tcgen05": (c_int, []),
    "dc_set_pdl": (c_int, [c_int]),
    "dc_get_pdl": (c_int, []),
    "dc_set_deterministic": (c_int, [c_int]),
    "dc_get_deterministic": (c_int, []),
    "dc_copy_view": (c_int, [dc_view, dc_view, c_void_p]),
    "dc_fill_zero": (c_int, [c_void_p, c_size_t, c_void_p]),
    "dc_ingest_hwc": (c_int, [c_void_p, ctypes.c_longlong, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "dc_i64_increment_many": (c_int, [c_void_p, c_int, c_void_p]),
    "dc_pack_weight": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dc_pack_weights_multi": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "dc_unpack_wgrad": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dc_conv_gemm_simt": (c_int, [POINTER(dc_conv_desc), dc_view, c_void_p, c_void_p, dc_view, c_void_p]),
    "dc_conv_wgrad_simt": (c_int, [POINTER(dc_conv_desc), dc_view, dc_view, c_void_p, c_void_p]),
    "dc_conv_wgrad_simt_ws_elems": (c_int64, [POINTER(dc_conv_desc), dc_view, dc_view]),
    "dc_conv_wgrad_simt_det": (c_int, [POINTER(dc_conv_desc), dc_view, dc_view, c_void_p, c_void_p, c_int64, c_void_p]),
    "dc_conv_gemm_tc": (c_int, [POINTER(dc_conv_desc), dc_view, c_void_p, c_void_p, dc_view, c_void_p]),
    "dc_conv_gemm_tc_bnstats": (c_int, [POINTER(dc_conv_desc), dc_view, c_void_p, c_void_p, dc_view, c_void_p, c_void_p]),
    "dc_conv_gemm_tc_halo_ok": (c_int, [POINTER(dc_conv_desc), dc_view, dc_view]),
    "dc_conv_gemm_tc_bn_eval": (c_int, [POINTER(dc_conv_desc), dc_view, c_void_p, c_void_p, dc_view, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_float, c_int, c_void_p]),
    "dc_conv_wgrad_tc": (c_int, [POINTER(dc_conv_desc), dc_view, dc_view, c_void_p, c_void_p]),
    "dc_conv_wgrad_tc_ws_elems": (c_int64, [POINTER(dc_conv_desc), dc_view, dc_view]),
    "dc_conv_wgrad_tc_det": (c_int, [POINTER(dc_conv_desc), dc_view, dc_view, c_void_p, c_void_p, c_int64, c_void_p]),
    "dc_dw_fwd": (c_int, [dc_view, c_void_p, c_int, c_int, dc_view, c_void_p]),
    "dc_dw_fwd_bn": (c_int, [POINTER(dc_bn_params), dc_view, c_void_p, dc_view, dc_view, c_void_p]),
    "dc_dw_bwd_data": (c_int, [dc_view, c_void_p, c_int, c_int, dc_view, c_int, c_void_p]),
    "dc_dw_bwd_weight": (c_int, [dc_view, dc_view, c_int, c_int, c_void_p, c_int, c_void_p]),
    "dc_dw_bwd_weight_ws_elems": (c_int64, [dc_view, dc_view, c_int, c_int]),
    "dc_dw_bwd_weight_det": (c_int, [dc_view, dc_view, c_int, c_int, c_void_p, c_int, c_void_p, c_int64, c_void_p]),
    "dc_dw_bwd_data_bnred": (c_int, [dc_view, c_void_p, dc_view, c_int, dc_view, dc_view, c_void_p, c_void_p, c_int, c_void_p]),
    "dc_bn_ws_bytes": (c_size_t, [c_int]),
    "dc_bn_stats": (c_int, [POINTER(dc_bn_params), dc_view, c_void_p]),
    "dc_bn_apply": (c_int, [POINTER(dc_bn_params), dc_view, dc_view, dc_view, c_void_p]),
    "dc_bn_bwd_reduce": (c_int, [POINTER(dc_bn_params), dc_view, dc_view, dc_view, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dc_bn_bwd_apply": (c_int, [POINTER(dc_bn_params), dc_view, dc_view, dc_view, c_void_p, dc_view, dc_view, c_void_p]),
    "dc_bn_bwd_apply_finalize": (c_int, [POINTER(dc_bn_params), dc_view, dc_view, dc_view, c_void_p, dc_view, dc_view, c_void_p,
                                         c_void_p, c_void_p]),
    "dc_bn_bwd_apply_reduced": (c_int, [POINTER(dc_bn_params), dc_view, dc_view, c_void_p, dc_view, dc_view, c_void_p, c_void_p,
                                        c_void_p]),
    "dc_bn_onepass_ok": (c_int, [c_int, c_int64, c_int, c_int]),
    "dc_bn_fwd_onepass": (c_int, [POINTER(dc_bn_params), dc_view, dc_view, dc_view, c_void_p]),
    "dc_bn_bwd_onepass": (c_int, [POINTER(dc_bn_params), dc_view, dc_view, dc_view, c_void_p, dc_view, dc_view, c_void_p,
                                  c_void_p, c_void_p]),
    "dc_channel_sum": (c_int, [dc_view, c_void_p, c_void_p, c_int, c_void_p]),
    "dc_gap_fwd": (c_int, [dc_view, c_void_p, c_void_p]),
    "dc_broadcast_hw": (c_int, [c_void_p, dc_view, c_void_p]),
    "dc_reduce_hw": (c_int, [dc_view, c_void_p, c_void_p]),
    "dc_gap_bwd": (c_int, [c_void_p, dc_view, c_int, c_void_p]),
    "dc_bilinear_fwd": (c_int, [dc_view, dc_view, c_void_p]),
    "dc_bilinear_bwd": (c_int, [dc_view, dc_view, c_int, c_void_p]),
    "dc_wce_fwd": (c_int, [dc_view, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dc_wce_bwd": (c_int, [dc_view, c_void_p, c_void_p, c_void_p, dc_view, c_void_p]),
    "dc_iou_counts": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "dc_argmax_iou": (c_int, [dc_view, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "dc_iou_finalize": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "dc_scale_f32": (c_int, [c_void_p, c_size_t, c_float, c_void_p]),
    "dc_adam_step_multi": (c_int, [c_void_p, c_int, c_int, c_double, c_double, c_double, c_double, c_double, c_double,
                                   c_double, c_int, c_void_p]),
    "dc_lars_step_multi": (c_int, [c_void_p, c_int, c_int, c_double, c_double, c_double, c_double, c_double, c_void_p, c_void_p]),
    "dc_lamb_step_multi": (c_int, [c_void_p, c_int, c_int, c_double, c_double, c_double, c_double, c_double, c_double,
                                   c_double, c_int, c_int, c_double, c_int, c_void_p, c_void_p]),
}

_LIB = None


def lib_path():
    return _build.LIB_PATH


def load():
    """Load (building first if necessary) the kernel library; raises if that is impossible."""
    global _LIB
    if _LIB is not None:
        return _LIB
    # always consult the source hash: a library built from older csrc/ must not be loaded silently (build() returns at once
    # when the stamp matches; it serialises concurrent ranks with a file lock)
    path = _build.build()
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = ABI mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.dc_abi_version() != DC_ABI_VERSION:
        raise RuntimeError("libdeepcam_b200.so: unexpected ABI version %d" % lib.dc_abi_version())
    _LIB = lib
    return lib


class KernelError(RuntimeError):
    pass


# kernels launched per successful C call (cudaMemsetAsync nodes are not counted as kernels)
_KERNELS_PER_CALL = {"dc_fill_zero": 0, "dc_wce_fwd": 2, "dc_channel_sum": 2, "dc_lamb_step_multi": 3, "dc_lars_step_multi": 2,
                     "dc_conv_wgrad_tc_det": 2, "dc_conv_wgrad_simt_det": 2, "dc_dw_bwd_weight_det": 2}
launch_count = 0          # number of deepcam_b200 kernels launched by this process (bench.py reports it)
launch_hist = {}


def check(rc, what=""):
    global launch_count
    n = _KERNELS_PER_CALL.get(what, 1)
    launch_count += n
    if n:
        launch_hist[what] = launch_hist.get(what, 0) + n
    if rc != 0:
        msg = load().dc_last_error_string().decode("utf-8", "replace")
        raise KernelError("%s failed (code %d): %s" % (what or "deepcam_b200 kernel", rc, msg))
