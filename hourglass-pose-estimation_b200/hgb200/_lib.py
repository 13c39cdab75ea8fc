"""ctypes binding of libhgb200.so.  Mirrors include/hg_api.h one to one."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libhgb200.so")


class HgError(RuntimeError):
    pass


def lib_path() -> str:
    return _LIB_PATH


class ConvDesc(C.Structure):
    """struct hg_conv_desc"""
    _fields_ = [
        ("in_", C.c_void_p), ("in2", C.c_void_p), ("weight", C.c_void_p), ("bias", C.c_void_p),
        ("in_scale", C.c_void_p), ("in_shift", C.c_void_p), ("residual", C.c_void_p), ("up_low", C.c_void_p),
        ("out", C.c_void_p), ("out_nchw_f32", C.c_void_p), ("err_word", C.c_void_p),
        ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("cin", C.c_int32), ("cin2", C.c_int32), ("cout", C.c_int32),
        ("ksize", C.c_int32), ("relu", C.c_int32), ("out_halo", C.c_int32),
        ("stats", C.c_void_p), ("pool_out", C.c_void_p), ("pool_in", C.c_void_p),
    ]


class PackEntry(C.Structure):
    """struct hg_pack_entry"""
    _fields_ = [
        ("src", C.c_void_p), ("src2", C.c_void_p), ("dst_f32", C.c_void_p), ("dst_fwd", C.c_void_p),
        ("dst_dgrad", C.c_void_p),
        ("co", C.c_int32), ("taps", C.c_int32), ("ci", C.c_int32),
        ("fwd_ld", C.c_int32), ("fwd_col0", C.c_int32), ("dgrad_ld", C.c_int32),
    ]


_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float

_SIGNATURES = {
    "hg_api_version": ([], C.c_int),
    "hg_last_error": ([C.c_char_p, C.c_size_t], C.c_size_t),
    "hg_check_device": ([], C.c_int),
    "hg_conv_nhwc_bf16": ([C.POINTER(ConvDesc), _vp], C.c_int),
    "hg_stem_im2col": ([_vp, _vp, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_halo_padded_elems": ([_i32, _i32, _i32, _i32], C.c_int64),
    "hg_conv3x3_halo_bf16": ([_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_conv3x3_k3_fusable": ([_i32, _i32, _i32], C.c_int),
    "hg_conv3x3_k3_fused_bf16": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp], C.c_int),
    "hg_stem_pack": ([_vp, _vp, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_stem_conv": ([_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp], C.c_int),
    "hg_maxpool2x2_nhwc": ([_vp, _vp, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_upsample2x_add_nhwc": ([_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_bn_relu_nhwc": ([_vp, _vp, _vp, _vp, _i64, _i32, _vp], C.c_int),
    "hg_nchw_f32_to_nhwc_bf16": ([_vp, _vp, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_nhwc_bf16_to_nchw_f32": ([_vp, _vp, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_decode_argmax": ([_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_decode_final_preds": ([_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_decode_final_preds_v2": ([_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_flip_average": ([_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_pck_dists": ([_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_joint_centers": ([_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_gaussian_target": ([_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_jmse_loss": ([C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32,
                      _f32, _vp], C.c_int),
    "hg_colreduce_scratch_bytes": ([_i64, _i32], C.c_int64),
    "hg_colstats_nhwc": ([_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp], C.c_int),
    "hg_bn_train_fwd": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _f32,
                         _vp], C.c_int),
    "hg_bn_bwd_reduce": ([_vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp], C.c_int),
    "hg_bn_bwd_apply": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_normalize_u8_nhwc": ([_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_preprocess_frames_u8": ([_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_dwconv3x3_nhwc": ([_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_dwconv3x3_wgrad": ([_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_maxpool2x2_bwd_nhwc": ([_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_sumpool2x2_nhwc": ([_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_add_inplace_bf16": ([_vp, _vp, _i64, _vp], C.c_int),
    "hg_nchw_f32_to_nhwc_bf16_pad": ([_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp], C.c_int),
    "hg_pack_weights": ([_vp, _i32, _i32, _vp], C.c_int),
    "hg_rmsprop_step": ([_vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _vp], C.c_int),
    "hg_small_gemm_f32": ([_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp], C.c_int),
    "hg_wgrad_bf16": ([_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp], C.c_int),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class _Lib:
    """Lazy loader: importing the package never needs the .so; the first op does, loudly."""

    def __init__(self):
        self._dll = None

    def _load(self):
        if self._dll is None:
            if not os.path.exists(_LIB_PATH):
                raise HgError(f"{_LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                              f"(or `make -C hourglass-pose-estimation_b200/csrc`). There is no CPU fallback.")
            dll = C.CDLL(_LIB_PATH)
            for name, (argtypes, restype) in _SIGNATURES.items():
                fn = getattr(dll, name)          # AttributeError if the library does not export it
                fn.argtypes = argtypes
                fn.restype = restype
            self._dll = dll
        return self._dll

    def __getattr__(self, name):
        return getattr(self._load(), name)

    def last_error(self) -> str:
        buf = C.create_string_buffer(512)
        self._load().hg_last_error(buf, 512)
        return buf.value.decode("utf-8", "replace")

    def check(self, rc: int, what: str = ""):
        if rc != 0:
            raise HgError(f"{what or 'libhgb200'} failed with code {rc}: {self.last_error()}")


lib = _Lib()
