"""Operator-level wrappers over the C ABI (used by the parity tests to exercise each kernel family in
isolation).  Host numpy in / out; activations are NHWC."""
import ctypes as C

import numpy as np

from . import _capi

_f = C.c_float


def resize_u8(src: np.ndarray, dw: int, dh: int, device: int = 0) -> np.ndarray:
    src = _capi.as_c(src, np.uint8)
    sh, sw = src.shape[:2]
    dst = np.empty((dh, dw, 3), np.uint8)
    _capi.check(_capi.lib().vsr_op_resize_u8(device, _capi.ptr(src, C.c_uint8), sh, sw, _capi.ptr(dst, C.c_uint8), dh, dw))
    return dst


def conv2d(x, weight, bias, ksize=3, dilation=1, lrelu=False, residual=None, device: int = 0) -> np.ndarray:
    x = _capi.as_c(x, np.float32)
    w = _capi.as_c(weight, np.float32)
    b = _capi.as_c(bias, np.float32)
    T, H, W, Cin = x.shape
    Cout = w.shape[0]
    out = np.empty((T, H, W, Cout), np.float32)
    res = _capi.as_c(residual, np.float32) if residual is not None else None
    flags = (1 if lrelu else 0) | (2 if res is not None else 0)
    _capi.check(_capi.lib().vsr_op_conv2d(device, _capi.ptr(x, _f), T, H, W, Cin, _capi.ptr(w, _f), _capi.ptr(b, _f), Cout, ksize,
                                          dilation, flags, _capi.ptr(res, _f) if res is not None else None, _capi.ptr(out, _f)))
    return out


def conv2d_s2(x, weight, bias, lrelu=False, device: int = 0) -> np.ndarray:
    x = _capi.as_c(x, np.float32)
    w = _capi.as_c(weight, np.float32)
    b = _capi.as_c(bias, np.float32)
    T, H, W, Cin = x.shape
    Cout = w.shape[0]
    out = np.empty((T, H // 2, W // 2, Cout), np.float32)
    _capi.check(_capi.lib().vsr_op_conv2d_s2(device, _capi.ptr(x, _f), T, H, W, Cin, _capi.ptr(w, _f), _capi.ptr(b, _f), Cout,
                                             1 if lrelu else 0, _capi.ptr(out, _f)))
    return out


def patch_attention(q, k, v, patches, device: int = 0) -> np.ndarray:
    q, k, v = (_capi.as_c(a, np.float32) for a in (q, k, v))
    T, H, W, Cc = q.shape
    pw = np.asarray([p[0] for p in patches], np.int32)
    ph = np.asarray([p[1] for p in patches], np.int32)
    out = np.empty_like(q)
    _capi.check(_capi.lib().vsr_op_patch_attention(device, _capi.ptr(q, _f), _capi.ptr(k, _f), _capi.ptr(v, _f), T, H, W, Cc,
                                                   len(patches), _capi.ptr(pw, C.c_int32), _capi.ptr(ph, C.c_int32),
                                                   _capi.ptr(out, _f)))
    return out


def upsample2x(x, device: int = 0) -> np.ndarray:
    x = _capi.as_c(x, np.float32)
    T, H, W, Cc = x.shape
    out = np.empty((T, 2 * H, 2 * W, Cc), np.float32)
    _capi.check(_capi.lib().vsr_op_upsample2x(device, _capi.ptr(x, _f), T, H, W, Cc, _capi.ptr(out, _f)))
    return out
