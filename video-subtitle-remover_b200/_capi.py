"""ctypes binding of include/vsr_b200.h.  No fallbacks: a missing library or a failing call raises."""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
CSRC = _HERE / "csrc"
LIB_PATH = Path(os.environ["VSR_B200_LIB"]) if os.environ.get("VSR_B200_LIB") else CSRC / "build" / "libvsr_b200.so"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]


class VsrError(RuntimeError):
    pass


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """nvcc cross-compiles the single translation unit csrc/engine.cu for sm_100a (no GPU needed)."""
    srcs = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [_HERE.parent / "include" / "vsr_b200.h"]
    if LIB_PATH.exists() and not force:
        newest = max(s.stat().st_mtime for s in srcs)
        if LIB_PATH.stat().st_mtime >= newest:
            return LIB_PATH
    LIB_PATH.parent.mkdir(parents=True, exist_ok=True)
    nvcc = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, *NVCC_FLAGS, str(CSRC / "engine.cu"), "-o", str(LIB_PATH), "-lcufft", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise VsrError("nvcc failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


class Config(C.Structure):
    _fields_ = [("model_w", C.c_int32), ("model_h", C.c_int32), ("n_patch", C.c_int32), ("patch_w", C.c_int32 * 4),
                ("patch_h", C.c_int32 * 4), ("neighbor_stride", C.c_int32), ("ref_length", C.c_int32), ("mode", C.c_int32)]


_lib = None

_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_pp = C.POINTER(C.c_void_p)

_PROTOS = {
    "vsr_last_error": (C.c_char_p, []),
    "vsr_version": (C.c_char_p, []),
    "vsr_device_count": (C.c_int, []),
    "vsr_sttn_default_config": (None, [C.POINTER(Config)]),
    "vsr_sttn_det_config": (None, [C.POINTER(Config)]),
    "vsr_sttn_create": (C.c_int, [_pp, C.c_int, C.POINTER(Config)]),
    "vsr_sttn_destroy": (None, [C.c_void_p]),
    "vsr_sttn_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, _f32p, _i64p, C.c_int]),
    "vsr_sttn_finalize_weights": (C.c_int, [C.c_void_p]),
    "vsr_sttn_inpaint_strip": (C.c_int, [C.c_void_p, _u8p, C.c_int, _f32p, _i32p]),
    "vsr_sttn_inpaint_strip_masked": (C.c_int, [C.c_void_p, _u8p, _u8p, C.c_int, _f32p, _i32p]),
    "vsr_sttn_inpaint_frames": (C.c_int, [C.c_void_p, _pp, C.c_int, C.c_int, C.c_int, _u8p, _pp]),
    "vsr_sttn_stage": (C.c_int, [C.c_void_p, _pp, C.c_int, C.c_int, C.c_int, _u8p]),
    "vsr_sttn_compute": (C.c_int, [C.c_void_p]),
    "vsr_sttn_fetch": (C.c_int, [C.c_void_p, _pp]),
    "vsr_sttn_submit": (C.c_int64, [C.c_void_p, _pp, C.c_int, C.c_int, C.c_int, _u8p]),
    "vsr_sttn_collect": (C.c_int, [C.c_void_p, C.c_int64, _pp]),
    "vsr_sttn_shard_begin": (C.c_int, [C.c_void_p, _pp, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, C.c_int, _pp, _i64p, _pp, _i64p]),
    "vsr_sttn_shard_windows": (C.c_int, [C.c_void_p]),
    "vsr_sttn_shard_finish": (C.c_int, [C.c_void_p, _pp]),
    "vsr_sttn_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "vsr_sttn_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "vsr_sttn_sync": (C.c_int, [C.c_void_p]),
    "vsr_sttn_stream": (C.c_void_p, [C.c_void_p]),
    "vsr_sttn_launch_count": (C.c_int64, [C.c_void_p]),
    "vsr_debug_tc_profile": (C.c_int, [C.POINTER(C.c_uint64), C.c_int]),
    "vsr_sttn_debug_read": (C.c_int, [C.c_void_p, C.c_char_p, _f32p, C.c_int64]),
    "vsr_sttn_profile": (C.c_int, [C.c_void_p, _f32p, _i64p, C.c_int]),
    "vsr_sttn_time_conv": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _f32p]),
    "vsr_create_mask": (C.c_int, [_u8p, C.c_int, C.c_int, _i32p, C.c_int, C.c_int]),
    "vsr_inpaint_area_by_mask": (C.c_int, [C.c_int, C.c_int, C.c_int, _u8p, C.c_int, _i32p, C.c_int]),
    "vsr_batch_sizes": (C.c_int, [C.c_int, C.c_int, _i32p, C.c_int]),
    "vsr_window_schedule": (C.c_int, [C.c_int, C.c_int, C.c_int, _i32p, _i32p, _i32p, C.c_int, C.c_int]),
    "vsr_rt_create": (C.c_int, [_pp, C.c_int]),
    "vsr_rt_destroy": (None, [C.c_void_p]),
    "vsr_rt_alloc": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_uint64)]),
    "vsr_rt_free": (C.c_int, [C.c_void_p, C.c_uint64]),
    "vsr_rt_upload": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int64]),
    "vsr_rt_download": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int64]),
    "vsr_rt_copy": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int64]),
    "vsr_rt_sync": (C.c_int, [C.c_void_p]),
    "vsr_rt_launch_count": (C.c_int64, [C.c_void_p]),
    "vsr_rt_scene_begin": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "vsr_rt_scene_frames": (C.c_int, [C.c_void_p, _pp, C.c_int, _i64p]),
    "vsr_rt_conv_create": (C.c_int, [C.c_void_p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_int, _i32p]),
    "vsr_rt_conv_create_split": (C.c_int, [C.c_void_p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "vsr_rt_conv": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_float,
                              C.c_float]),
    "vsr_rt_conv_ex": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_float,
                                 C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vsr_rt_pad": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vsr_rt_zero_upsample2x": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64]),
    "vsr_rt_add_slices": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int64,
                                    C.c_float, C.c_float]),
    "vsr_rt_hswish_affine": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int64, C.c_float, C.c_float, C.c_float]),
    "vsr_rt_se_create": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.POINTER(C.c_int)]),
    "vsr_rt_se_gate": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_int64, C.c_int, C.c_float, C.c_uint64]),
    "vsr_rt_pp_frames": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_uint64]),
    "vsr_rt_instnorm": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_uint64]),
    "vsr_rt_context_split": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.c_uint64, C.c_int, C.c_uint64, C.c_int]),
    "vsr_rt_corr_volume": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int]),
    "vsr_rt_corr_pool": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int]),
    "vsr_rt_corr_lookup": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), _i32p, _i32p, _i32p, C.c_uint64, C.c_int, C.c_int, C.c_int64, C.c_uint64, C.c_int]),
    "vsr_rt_gru_rh": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.c_int64]),
    "vsr_rt_gru_update": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.c_int64]),
    "vsr_rt_flow_update": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int64, C.c_int]),
    "vsr_rt_convex_upsample": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64]),
    "vsr_rt_img_prop_step": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_uint64]),
    "vsr_rt_prop_state": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint64]),
    "vsr_rt_rfc_input": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64]),
    "vsr_rt_pad_replicate": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vsr_rt_leaky_relu": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.c_float]),
    "vsr_rt_temporal_taps": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int64, C.c_int, C.c_uint64, C.c_int]),
    "vsr_rt_deform_cols": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_float, C.c_uint64,
                                     C.c_int, C.c_int, C.c_int64, C.c_uint64, C.c_int]),
    "vsr_rt_rfc_combine": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64]),
    "vsr_rt_upsample2x_bilinear": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64]),
    "vsr_rt_gen_input": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint64]),
    "vsr_rt_flow_down4": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint64]),
    "vsr_rt_prop_masks": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint64]),
    "vsr_rt_featprop_cond": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_int]),
    "vsr_rt_write_extra": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int64]),
    "vsr_rt_unfold7s3": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int]),
    "vsr_rt_fold7s3": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64]),
    "vsr_rt_layernorm": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64]),
    "vsr_rt_pool4": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64]),
    "vsr_rt_window_attention": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64]),
    "vsr_rt_pred_to_rgb8": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int64, _u8p]),
    "vsr_rt_residual_add": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int64, C.c_int]),
    "vsr_rt_fft_r2c": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64]),
    "vsr_rt_fft_c2r": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int]),
    "vsr_rt_lama_input": (C.c_int, [C.c_void_p, _u8p, _u8p, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vsr_rt_lama_output": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, _u8p]),
    "vsr_rt_absmax": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.POINTER(C.c_float)]),
    "vsr_rt_download_channel": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_float, _f32p]),
    "vsr_rt_overflow": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "vsr_rt_capture_begin": (C.c_int, [C.c_void_p]),
    "vsr_rt_capture_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "vsr_rt_graph_launch": (C.c_int, [C.c_void_p, C.c_int]),
    "vsr_rt_graph_destroy": (C.c_int, [C.c_void_p, C.c_int]),
    "vsr_rt_elementwise": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int64, C.c_int, C.c_uint64, C.c_uint64,
                                     C.c_float, C.c_float]),
    "vsr_rt_upsample_nearest": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int]),
    "vsr_rt_maxpool2x2s1": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64]),
    "vsr_rt_copy_channels": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int64]),
    "vsr_rt_det_preprocess": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int]),
    "vsr_op_resize_u8": (C.c_int, [C.c_int, _u8p, C.c_int, C.c_int, _u8p, C.c_int, C.c_int]),
    "vsr_op_conv2d": (C.c_int, [C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_int, C.c_int, C.c_int,
                                C.c_int, _f32p, _f32p]),
    "vsr_op_conv2d_s2": (C.c_int, [C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_int, C.c_int, _f32p]),
    "vsr_op_patch_attention": (C.c_int, [C.c_int, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p,
                                         _f32p]),
    "vsr_op_upsample2x": (C.c_int, [C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]),
}

EXPORTED_SYMBOLS = tuple(_PROTOS)


def lib():
    """Load (once) the C-ABI library.  Raises VsrError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise VsrError(f"{LIB_PATH} is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(vsr_b200 has no CPU fallback)")
    try:
        L = C.CDLL(str(LIB_PATH))
    except OSError as e:  # e.g. libcudart/libcuda unresolved
        raise VsrError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in _PROTOS.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class VsrRangeError(VsrError):
    """VSR_ERR_RANGE: the job must be repeated with the engine option attn_direct = 0"""


def check(rc: int):
    if rc == -5:
        raise VsrRangeError(lib().vsr_last_error().decode("utf-8", "replace") + " (code -5)")
    if rc < 0:
        raise VsrError(lib().vsr_last_error().decode("utf-8", "replace") + f" (code {rc})")
    return rc


def ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def as_c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)
