"""ctypes binding of libctl_b200.so (C ABI: include/ctl_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C csrc`.  There is no
fallback of any kind: if the shared object is missing, or a compute entry point is called
without a CUDA device, the call raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libctl_b200.so")

CTL_OK, CTL_ERR_INVALID, CTL_ERR_INDEX, CTL_ERR_UNSUPPORTED, CTL_ERR_CUDA = 0, 1, 2, 3, 4
CTL_F32, CTL_BF16 = 0, 1
MODE_CHANNEL, MODE_SPATIAL = 0, 1
ACT_NONE, ACT_LRELU, ACT_RELU, ACT_SIGMOID = 0, 1, 2, 3

_c = ctypes
_vp, _i, _i64, _u64, _f = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_uint64, _c.c_float

# name -> (restype, argtypes); mirrors include/ctl_b200.h one to one
SIGNATURES = {
    "ctl_version": (_i, []),
    "ctl_last_error": (_c.c_char_p, []),
    "ctl_device_sm_count": (_i, []),
    "ctl_saliency_reduce": (_i, [_vp, _i, _i64, _i64, _i64, _i, _vp, _vp]),
    "ctl_topp_mask_apply": (_i, [_vp, _vp, _i, _i64, _i64, _i64, _i, _i64, _i, _vp, _u64, _u64, _i64,
                                 _vp, _vp, _vp, _i, _vp]),
    "ctl_saliency_mask_apply": (_i, [_vp, _i, _vp, _i, _i64, _i64, _i64, _i, _i64, _i, _vp, _u64, _u64, _i64,
                                     _vp, _vp, _vp, _vp, _i, _vp]),
    "ctl_channel_dropout": (_i, [_vp, _i, _i64, _i64, _i64, _f, _f, _vp, _u64, _u64, _i64, _vp, _i, _vp, _vp, _vp]),
    "ctl_saliency_mask_apply_dyn": (_i, [_vp, _i, _vp, _i, _i64, _i64, _i64, _i, _i, _vp, _u64, _vp,
                                         _vp, _vp, _vp, _vp, _i, _vp]),
    "ctl_channel_dropout_dyn": (_i, [_vp, _i, _i64, _i64, _i64, _f, _f, _u64, _vp, _vp, _i, _vp, _vp, _vp]),
    "ctl_philox_uniform": (_i, [_u64, _u64, _u64, _i64, _vp, _vp]),
    "ctl_conv2d_n_tile": (_i, [_i, _i, _i]),
    "ctl_conv2d_vpacked": (_i, [_i, _i, _i, _i]),
    "ctl_pack_conv_weight": (_i, [_vp, _i64, _i64, _i, _i, _i, _vp, _vp]),
    "ctl_pack_conv_weights_batched": (_i, [_vp, _i64, _i64, _vp]),
    "ctl_conv2d_c8_bf16": (_i, [_vp, _i64, _i64, _i64, _i64, _vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp,
                                _vp, _vp]),
    "ctl_bn_affine_from_sums": (_i, [_vp, _i64, _i64, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp]),
    "ctl_nchw_to_c8": (_i, [_vp, _i, _i64, _i64, _i64, _i64, _vp, _vp]),
    "ctl_c8_to_nchw": (_i, [_vp, _i64, _i64, _i64, _i64, _vp, _i, _vp]),
    "ctl_stem_conv3x3_c8": (_i, [_vp, _vp, _i, _f, _i64, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _i, _vp, _vp]),
    "ctl_head_conv1x1_c8": (_i, [_vp, _i64, _i64, _i64, _i64, _vp, _vp, _i64, _i, _vp, _vp]),
    "ctl_upsample2x_c8": (_i, [_vp, _i64, _i64, _i64, _i64, _vp, _vp]),
    "ctl_bn_workspace_bytes": (_c.c_size_t, [_i64, _i64]),
    "ctl_bn_batch_affine_c8": (_i, [_vp, _i64, _i64, _i64, _i64, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f,
                                    _vp]),
    "ctl_scale_shift_act_c8": (_i, [_vp, _i64, _i64, _i64, _i64, _vp, _vp, _i, _vp, _vp]),
    "ctl_scale_shift_upadd_act_c8": (_i, [_vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _i, _vp, _vp]),
    "ctl_conv_wgrad_c8_bf16": (_i, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _i, _vp, _i64, _i64, _i64, _vp]),
    "ctl_reduce_workspace_bytes": (_c.c_size_t, [_i64, _i64]),
    "ctl_channel_sums_c8": (_i, [_vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "ctl_bn_bwd_reduce_c8": (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp,
                                  _vp, _vp, _vp]),
    "ctl_bn_bwd_apply_c8": (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i, _vp, _vp, _vp, _vp, _vp]),
    "ctl_bn_bwd_c8": (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                           _vp]),
    "ctl_conv2d_c8_bf16_bnbwd": (_i, [_vp, _i64, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "ctl_bn_bwd_apply_totals_c8": (_i, [_vp, _vp, _i64, _i64, _i64, _i64, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                        _vp]),
    "ctl_bn_apply_from_sums_c8": (_i, [_vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _f, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp,
                                       _vp, _f, _vp]),
    "ctl_act_bwd_c8": (_i, [_vp, _vp, _i64, _i64, _i64, _i64, _i, _vp, _vp]),
    "ctl_downsample2x_sum_c8": (_i, [_vp, _i64, _i64, _i64, _i64, _vp, _vp]),
    "ctl_zero_stuff2x_c8": (_i, [_vp, _i64, _i64, _i64, _i64, _vp, _vp]),
    "ctl_split_parity2x2_c8": (_i, [_vp, _i64, _i64, _i64, _i64, _vp, _vp]),
    "ctl_head_bwd_c8": (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _i64, _i, _vp, _vp, _vp, _vp]),
    "ctl_stem_input_c8": (_i, [_vp, _vp, _i, _f, _i64, _i64, _i64, _i64, _vp, _vp]),
    "ctl_stem_wgrad_c8": (_i, [_vp, _vp, _vp, _i, _f, _i64, _i64, _i64, _i64, _vp, _vp]),
    "ctl_stem_dgrad_c8": (_i, [_vp, _vp, _i, _f, _i64, _i64, _i64, _i64, _vp, _vp, _vp]),
    "ctl_ce2d_fwd": (_i, [_vp, _vp, _i64, _i64, _i64, _i64, _c.c_double, _vp, _vp, _vp]),
    "ctl_ce2d_bwd": (_i, [_vp, _vp, _i64, _i64, _i64, _i64, _c.c_double, _vp, _vp, _vp]),
    "ctl_adam_flat": (_i, [_vp, _vp, _vp, _vp, _c.POINTER(_i64), _i, _c.c_uint, _vp, _c.c_double, _c.c_double, _c.c_double,
                           _c.c_double, _c.c_double, _c.c_double, _i, _vp]),
    "ctl_sse_fwd": (_i, [_vp, _vp, _i64, _c.c_double, _vp, _vp, _vp]),
    "ctl_sse_bwd": (_i, [_vp, _vp, _i64, _c.c_double, _vp, _vp, _vp]),
    "ctl_confusion_update": (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "ctl_confusion_scores": (_i, [_vp, _i64, _vp, _vp]),
    "ctl_conv2d_c8_bf16_saliency": (_i, [_vp, _i64, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _i, _i, _vp]),
    "ctl_saliency_sums_mask_apply": (_i, [_vp, _vp, _i64, _i64, _i64, _i, _i64, _i, _vp, _u64, _u64, _i64, _vp, _vp, _vp,
                                          _vp, _vp, _vp, _vp]),
}

_lib = None
# kernels launched through this binding since import (bench.py reports the count inside its timed
# region as `gpu_launches`); name -> kernels per successful call
KERNELS_PER_CALL = {"ctl_saliency_reduce": 1, "ctl_topp_mask_apply": 2, "ctl_saliency_mask_apply": 3,
                    "ctl_channel_dropout": 1, "ctl_saliency_mask_apply_dyn": 3, "ctl_channel_dropout_dyn": 1,
                    "ctl_philox_uniform": 1, "ctl_conv2d_c8_bf16": 1,
                    "ctl_nchw_to_c8": 1, "ctl_c8_to_nchw": 1, "ctl_stem_conv3x3_c8": 1, "ctl_head_conv1x1_c8": 1,
                    "ctl_upsample2x_c8": 1, "ctl_bn_batch_affine_c8": 2, "ctl_scale_shift_act_c8": 1,
                    "ctl_conv_wgrad_c8_bf16": 1, "ctl_channel_sums_c8": 2, "ctl_bn_affine_from_sums": 1, "ctl_pack_conv_weight": 1, "ctl_bn_bwd_reduce_c8": 2,
                    "ctl_bn_bwd_apply_c8": 1, "ctl_act_bwd_c8": 1, "ctl_downsample2x_sum_c8": 1,
                    "ctl_zero_stuff2x_c8": 1, "ctl_split_parity2x2_c8": 1, "ctl_head_bwd_c8": 1,
                    "ctl_stem_wgrad_c8": 1, "ctl_stem_dgrad_c8": 1, "ctl_ce2d_fwd": 1, "ctl_ce2d_bwd": 1, "ctl_scale_shift_upadd_act_c8": 1, "ctl_pack_conv_weights_batched": 1, "ctl_stem_input_c8": 1,
                    "ctl_adam_flat": 2, "ctl_sse_fwd": 1, "ctl_sse_bwd": 1, "ctl_confusion_update": 1,
                    "ctl_confusion_scores": 1, "ctl_conv2d_c8_bf16_saliency": 1, "ctl_saliency_sums_mask_apply": 1,
                    "ctl_bn_bwd_c8": 2, "ctl_conv2d_c8_bf16_bnbwd": 1, "ctl_bn_bwd_apply_totals_c8": 1, "ctl_bn_apply_from_sums_c8": 1}
LAUNCHES = {"count": 0}


class CtlError(RuntimeError):
    pass


def load():
    """Loads the shared object once; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CtlError(
            "libctl_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C %s/csrc`.  This package has no CPU / PyTorch fallback." % (LIB_PATH, _HERE))
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = _Counted(lib)
    return _lib


class _Counted:
    """Thin proxy that counts the kernels each successful compute call launched."""

    def __init__(self, lib):
        self._lib = lib
        for name in SIGNATURES:
            fn = getattr(lib, name)
            per = KERNELS_PER_CALL.get(name)
            setattr(self, name, self._wrap(fn, per) if per else fn)

    @staticmethod
    def _wrap(fn, per):
        def call(*args):
            rc = fn(*args)
            if rc == CTL_OK:
                LAUNCHES["count"] += per
            return rc
        return call


def last_error():
    msg = load().ctl_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc):
    """Maps a ctl_status to the exception the reference would raise at that point."""
    if rc == CTL_OK:
        return
    msg = last_error()
    if rc == CTL_ERR_INDEX:
        raise IndexError(msg)
    if rc == CTL_ERR_INVALID:
        raise ValueError(msg)
    if rc == CTL_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise CtlError("ctl_b200 CUDA failure: " + msg)
