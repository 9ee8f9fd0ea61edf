"""Operator-level parity cases of the ProPainter path, shared by two suites:

  tests/test_pp_abi_emulated.py  (CPU)  the kernels' and entry points' real source compiled for the host (tests/emu/), see that file;
  tests/test_gpu_pp_ops.py       (GPU)  the same cases on the device through the shipped library — the bring-up suite of DESIGN.md §7.

Every case runs one operator twice: on the numpy stand-in of the runtime (tests/fake_rt.py, the thing the CPU pipeline tests are built on) and
through the product's wrapper class -> ctypes prototypes -> `vsr_rt_*` entry point -> kernel, and compares the two."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from fake_rt import FakeRuntime

EMU = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "video-subtitle-remover_b200", "csrc")


def load_emu_library():
    """Build (when stale) and load the host build of the ProPainter entry points: tests/emu/make_abi_emu.py cuts them out of csrc/engine.cu,
    g++ compiles them with csrc/pp_ops.cuh over tests/emu/cuda_emu.h; the alignment sanitizer aborts on a misaligned 16-byte access.  The
    shipped ctypes prototypes (vsr_b200/_capi.py) are applied to it."""
    from vsr_b200 import _capi

    build = os.path.join(EMU, "build")
    out, gen = os.path.join(build, "libabi_emu.so"), os.path.join(build, "abi_emu.cpp")
    srcs = [os.path.join(EMU, f) for f in ("make_abi_emu.py", "abi_prelude.h", "cuda_emu.h")] + [os.path.join(CSRC, "engine.cu"), os.path.join(CSRC, "pp_ops.cuh"),
                                                                                                os.path.join(ROOT, "include", "vsr_b200.h")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(f) for f in srcs):
        os.makedirs(build, exist_ok=True)
        subprocess.run([sys.executable, srcs[0], os.path.join(CSRC, "engine.cu"), gen], check=True)
        subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-fsanitize=alignment", "-fno-sanitize-recover=alignment", "-I",
                        os.path.join(EMU, "stubs"), "-I", EMU, gen, "-o", out], check=True)
    L = C.CDLL(out)
    for name, (res, args) in _capi._PROTOS.items():
        if hasattr(L, name):
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
    L.emu_rt_create.restype = C.c_void_p
    L.emu_launches.restype, L.emu_launches.argtypes = C.c_long, [C.c_void_p]
    return L


def bind_wrapper(lib):
    """the product's wrapper class (_GenRuntime and its bases) bound to the host build of the C ABI"""
    from vsr_b200.propainter_generator import _GenRuntime

    rt = object.__new__(_GenRuntime)
    rt.L, rt.h = lib, C.c_void_p(lib.emu_rt_create())
    return rt


class HostBackend:
    """the host build of the entry points (tests/emu): "device memory" is numpy memory, pointers are host addresses"""

    def __init__(self, lib):
        self.lib = lib

    def runtime(self):
        return bind_wrapper(self.lib)

    def launches(self, rt):
        return self.lib.emu_launches(rt.h)

    def place(self, rt, host):
        return host.ctypes.data                                  # zero-copy: the kernels work on the numpy buffer itself

    def upload(self, rt, ptr, host):
        pass

    def download(self, rt, ptr, host):
        pass


class DeviceBackend:
    """the shipped library on cuda:0"""

    def runtime(self):
        from vsr_b200.propainter_generator import _GenRuntime

        return _GenRuntime("cuda:0")

    def launches(self, rt):
        return rt.launch_count

    def place(self, rt, host):
        p = rt.alloc(max(host.nbytes, 16))
        self.upload(rt, p, host)
        return p

    def upload(self, rt, ptr, host):
        from vsr_b200 import _capi

        _capi.check(rt.L.vsr_rt_upload(rt.h, ptr, host.ctypes.data_as(C.c_void_p), host.nbytes))

    def download(self, rt, ptr, host):
        from vsr_b200 import _capi

        _capi.check(rt.L.vsr_rt_download(rt.h, ptr, host.ctypes.data_as(C.c_void_p), host.nbytes))


class Pair:
    """one tensor in both worlds: `fake` lives in the stand-in's slot memory, `real` is fp16 memory behind a product _Tensor"""

    def __init__(self, dual, n, h, w, cp):
        from vsr_b200.dbnet import _Tensor

        self.dual, self.shape = dual, (n, h, w, cp)
        self.fake = _Tensor(dual.fake.alloc(n * h * w * cp * 2), cp, h, w, cp, n=n)
        self.bits = np.zeros((n, h, w, cp), np.uint16)
        self.real = _Tensor(dual.backend.place(dual.real, self.bits), cp, h, w, cp, n=n)

    @property
    def view(self):
        return self.dual.fake._v4(self.fake)

    def push(self):
        """stand-in contents -> rounded to fp16 -> the other world (after a test edited the stand-in's values)"""
        self.view[:] = self.view.astype(np.float16).astype(np.float32)
        self.bits[:] = self.view.astype(np.float16).view(np.uint16)
        self.dual.backend.upload(self.dual.real, self.real.ptr, self.bits)
        return self

    def check(self, what, ulps=2, atol=0.0, channels=None):
        self.dual.backend.download(self.dual.real, self.real.ptr, self.bits)
        got = self.bits.view(np.float16).astype(np.float32)
        want = self.view.astype(np.float16).astype(np.float32)
        if channels is not None:
            got, want = got[..., :channels], want[..., :channels]
        tol = ulps * np.maximum(np.abs(want), 2.0 ** -14) * 2.0 ** -10 + atol      # atol: sums with cancellation (fp32 summation order differs)
        bad = np.abs(got - want) > tol
        assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.size} differ, worst {np.abs(got - want).max():.4g}"


class Ptr:
    """a Pair handed to an operator as a bare pointer (the correlation pyramid's buffers)"""

    def __init__(self, pair):
        self.fake, self.real = pair.fake.ptr, pair.real.ptr


class Raw:
    """a side buffer (fp32 flow, u8 mask, int32 index list) in both worlds; operators take it as a bare device pointer"""

    def __init__(self, dual, arr, contiguous_f32=False):
        self.dual = dual
        self.host = np.ascontiguousarray(arr).copy()
        # the stand-in keeps fp32 data one value per two slots (byte offsets stay valid) except small parameter vectors (upload_f32)
        self.fake = dual.fake.upload_f32(self.host) if contiguous_f32 else dual.fake.upload_bytes(self.host)
        self.real = dual.backend.place(dual.real, self.host)

    def pull(self):
        self.dual.backend.download(self.dual.real, self.real, self.host)
        return self.host

    def fake_f32(self):
        return self.dual.fake._raw32(self.fake, self.host.size).reshape(self.host.shape)


class Dual:
    def __init__(self, backend, seed):
        self.backend = backend
        self.fake = FakeRuntime()
        self.real = backend.runtime()
        self.rng = np.random.default_rng(seed)

    def close(self):
        if isinstance(self.backend, DeviceBackend):
            self.real.close()

    def tensor(self, n, h, w, cp, scale=1.0, fill=True):
        p = Pair(self, n, h, w, cp)
        if fill:
            p.view[:] = self.rng.standard_normal((n, h, w, cp)) * scale
            p.push()
        return p

    def f32(self, arr):
        return Raw(self, np.asarray(arr, np.float32))

    def param(self, arr):
        return Raw(self, np.asarray(arr, np.float32), contiguous_f32=True)

    def u8(self, arr):
        return Raw(self, np.asarray(arr, np.uint8))

    def ints(self, arr):
        return Raw(self, np.asarray(arr, np.int32))

    def call(self, op, *args, **kw):
        """the same operator through the stand-in and through wrapper -> C ABI -> kernel; returns both results"""
        def side(a, which):
            if isinstance(a, (Pair, Raw, Ptr)):
                return getattr(a, which)
            if isinstance(a, (list, tuple)) and a and isinstance(a[0], tuple):       # correlation pyramid: [(Pair, H, W, pitch)]
                return [(getattr(t[0], which).ptr,) + tuple(t[1:]) for t in a]
            return a

        before = self.backend.launches(self.real)
        want = getattr(self.fake, op)(*[side(a, "fake") for a in args], **{k: side(v, "fake") for k, v in kw.items()})
        got = getattr(self.real, op)(*[side(a, "real") for a in args], **{k: side(v, "real") for k, v in kw.items()})
        assert self.backend.launches(self.real) > before
        return want, got


def case_frames_and_states(make):
    d = make(0)
    T, H, W = 2, 6, 10
    frames = [d.rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(T)]
    y = d.tensor(T, H, W, 8, fill=False)
    d.call("frames", frames, y)
    y.check("frames", ulps=0)
    mask = d.u8((d.rng.random((H, W)) > 0.6) * 255)
    st = d.tensor(T, H, W, 8, fill=False)
    d.call("prop_state", y, mask, None, st)
    st.check("prop_state (init)", ulps=0)
    prop, st2 = d.tensor(T, H, W, 8), d.tensor(T, H, W, 8, fill=False)
    d.call("prop_state", y, mask, prop, st2)
    st2.check("prop_state (compose)", ulps=0)


def case_instnorm(make, relu, cp=64):
    d = make(1)
    x = d.tensor(2, 9, 13, cp, scale=3.0)
    x.view[:] += np.linspace(-2, 2, cp, dtype=np.float32)
    x.push()
    y = d.tensor(2, 9, 13, cp, fill=False)
    d.call("instnorm", x, y, relu)
    y.check("instnorm", ulps=3, atol=1e-3)


def case_context_split_and_gru(make):
    d = make(2)
    N, H, W = 1, 5, 7
    x, net, inp = d.tensor(N, H, W, 256), d.tensor(N, H, W, 384), d.tensor(N, H, W, 384)
    d.call("context_split", x, net, inp)
    net.check("context_split.net")
    inp.check("context_split.inp")
    r, hsrc, out = d.tensor(N, H, W, 128, 2.0), d.tensor(N, H, W, 384), d.tensor(N, H, W, 384)
    d.call("gru_rh", r, hsrc, out)
    out.check("gru_rh")
    z, q, hio = d.tensor(N, H, W, 128, 2.0), d.tensor(N, H, W, 128, 2.0), d.tensor(N, H, W, 384)
    d.call("gru_update", z, q, hio)
    hio.check("gru_update")


def case_corr_pool_and_lookup(make, hh=8, ww=12):
    d = make(3)                         # hh x ww: the 1/8-resolution map; level l has (hh >> l, ww >> l) targets per source pixel (odd sizes drop a row / column)
    px = hh * ww
    H, W = hh, ww
    pitch = (H * W + 7) // 8 * 8
    levels = [(d.tensor(1, 1, px, pitch), H, W, pitch)]
    for _ in range(3):
        oh, ow = H // 2, W // 2
        op = (oh * ow + 7) // 8 * 8
        dst = d.tensor(1, 1, px, op, fill=False)
        src = levels[-1][0]
        d.call("corr_pool", Ptr(src), px, H, W, pitch, Ptr(dst), op)
        dst.check(f"corr_pool {H}x{W}", channels=oh * ow)
        dst.push()
        levels.append((dst, oh, ow, op))
        H, W, pitch = oh, ow, op
    flow = d.f32(d.rng.standard_normal((px, 2)) * 3)
    out = d.tensor(1, hh, ww, 384, fill=False)
    d.call("corr_lookup", levels, flow, hh, ww, px, out)
    out.check("corr_lookup", ulps=3, atol=1e-3, channels=324)
    assert np.abs(out.view[..., :324]).max() > 0.5


def case_flow_update(make, add):
    d = make(4)
    N, H, W = 2, 4, 6
    f = d.f32(d.rng.standard_normal((N * H * W, 2)) * 4)
    delta, f16, a, c = d.tensor(N, H, W, 64), d.tensor(N, H, W, 8), d.tensor(N, H, W, 384), d.tensor(N, H, W, 384)
    d.call("flow_update", f, delta, f16, a, c, 382, add)
    np.testing.assert_array_equal(f.pull(), f.fake_f32())
    for t, name in ((f16, "flow16"), (a, "dst_a"), (c, "dst_b")):
        t.check("flow_update." + name, ulps=0)
    d.call("flow_update", f, None, f16, None, None, 0, 0)        # the initial refresh: no delta, no GRU tensors
    f16.check("flow_update without destinations", ulps=0)


def case_convex_upsample(make):
    d = make(5)
    N, h, w = 2, 4, 5
    f = d.f32(d.rng.standard_normal((N * h * w, 2)) * 3)
    mask = d.tensor(N, h, w, 576, 2.0)
    out = d.f32(np.zeros(N * 2 * 64 * h * w))
    d.call("convex_upsample", f, mask, N, h, w, out)
    np.testing.assert_allclose(out.pull(), out.fake_f32(), rtol=2e-5, atol=2e-5)
    assert np.abs(out.host).max() > 1


def case_img_prop_step(make, H=12, W=20):
    d = make(6)
    prev, cur = d.tensor(1, H, W, 8), d.tensor(1, H, W, 8)
    prev.view[..., 3] = d.rng.random((1, H, W)) > 0.93       # few holes left in the source frame, half of the current frame missing
    cur.view[..., 3] = d.rng.random((1, H, W)) > 0.5
    prev.push(), cur.push()
    fp = d.f32(d.rng.standard_normal((2, H, W)) * 2)
    fc = d.f32(-fp.host + d.rng.standard_normal((2, H, W)).astype(np.float32) * 0.4)        # consistent for some pixels, not for others
    out = d.tensor(1, H, W, 8, fill=False)
    before = cur.view[..., 3].copy()
    d.call("img_prop_step", prev, cur, fp, fc, out)
    out.check("img_prop_step", ulps=0)
    assert 0.02 < (out.view[..., 3] != before).mean() < 0.98                      # both branches of the fill decision were taken


def case_rfc_input_combine(make, reverse):
    d = make(7)
    N, H, W = 3, 5, 6
    fl = d.f32(d.rng.standard_normal((N, 2, H, W)) * 3)
    m = d.u8((d.rng.random((H, W)) > 0.5) * 255)
    out = d.tensor(N, H, W, 8)
    d.call("rfc_input", fl, m, N, H, W, reverse, out)
    out.check("rfc_input", ulps=0)
    pred = d.tensor(N, H, W, 64)
    o = d.f32(np.zeros((N, 2, H, W)))
    d.call("rfc_combine", pred, fl, m, N, H, W, reverse, o)
    np.testing.assert_array_equal(o.pull(), o.fake_f32())


def case_pad_leaky_taps_extra(make):
    d = make(8)
    T, H, W, cp = 3, 5, 7, 16
    x = d.tensor(T, H, W, cp)
    y = d.tensor(T, H + 3, W + 4, cp, fill=False)
    d.call("pad_replicate", x, y, 1, 2)
    y.check("pad_replicate", ulps=0)
    d.call("leaky", x, 0.2)
    x.check("leaky", ulps=1)
    x.push()
    taps = d.tensor(T, H, W, 64)
    d.call("temporal_taps", x, taps)
    taps.check("temporal_taps", ulps=0, channels=3 * cp)
    src8, dst = d.tensor(T, H, W, 8), d.tensor(T, H, W, 64)
    d.call("write_extra", src8, dst, 37, 5)
    dst.check("write_extra", ulps=0)


def case_deform_cols(make, two_inputs, with_flow, G=4, Cc=32):
    d = make(9)
    n, H, W = 2, 6, 7                            # 8 channels per group (Cc / G): the 16-byte path; otherwise the scalar path
    xa = d.tensor(n, H, W, max(64, Cc))
    xb = d.tensor(n, H, W, max(64, Cc)) if two_inputs else None
    om = d.tensor(n, H, W, max(128, (27 * G + 7) // 8 * 8), 0.7)
    cols = d.tensor(n, H, W, 9 * Cc + 32)
    fl = d.f32(d.rng.standard_normal((n * H * W, 2)) * 2) if with_flow else 0
    d.call("deform_cols", xa, 16 if two_inputs else Cc, xb, Cc, G, om, 3.0, fl, cols)
    cols.check("deform_cols", ulps=4, atol=1e-3)


def case_gen_input_flow_down_masks(make):
    d = make(10)
    T, H, W = 4, 8, 12
    state = d.tensor(T, H, W, 8)
    m = d.u8((d.rng.random((H, W)) > 0.5) * 255)
    ids = d.ints([2, 0, 3])
    gin = d.tensor(3, H, W, 8)
    d.call("gen_input", state, m, ids, 3, gin)
    gin.check("gen_input", ulps=0)
    fl = d.f32(d.rng.standard_normal((T, 2, H, W)) * 3)
    o = d.f32(np.zeros(3 * (H // 4) * (W // 4) * 2))
    d.call("flow_down4", fl, ids, 3, H, W, o)
    np.testing.assert_allclose(o.pull(), o.fake_f32(), rtol=1e-6, atol=1e-6)
    gin.push()
    pm = d.tensor(3, H // 4, W // 4, 8)
    d.call("prop_masks", gin, pm)
    pm.check("prop_masks", ulps=0)


def case_featprop_cond(make):
    d = make(11)
    H, W, Cc = 7, 9, 16
    prop, cur, masks = d.tensor(1, H, W, Cc), d.tensor(1, H, W, Cc), d.tensor(1, H, W, 8)
    fp = d.f32(d.rng.standard_normal((H * W, 2)) * 2)
    fc = d.f32(-fp.host + d.rng.standard_normal((H * W, 2)).astype(np.float32) * 0.4)
    cond = d.tensor(1, H, W, 64)
    d.call("featprop_cond", prop, cur, fp, fc, masks, cond)
    cond.check("featprop_cond", ulps=3, atol=1e-3)
    assert 0.05 < cond.view[..., 2 * Cc + 2].mean() < 0.95


def case_unfold_fold(make, gelu, h=10, w=13):
    d = make(12)
    n, Cc = 2, 8
    fh, fw = (h + 6 - 7) // 3 + 1, (w + 6 - 7) // 3 + 1
    x = d.tensor(n, h, w, Cc)
    tok = d.tensor(n, fh, fw, 49 * Cc + 8)
    d.call("unfold7s3", x, tok, gelu)
    tok.check("unfold7s3", ulps=2, atol=2e-4 if gelu else 0.0)      # 1 + erf(x / sqrt 2) cancels for x < -3: erff implementations differ there
    tok.push()
    for norm in (False, True):
        out = d.tensor(n, h, w, Cc, fill=False)
        d.call("fold7s3", tok, out, Cc, norm)
        out.check(f"fold7s3 normalise={norm}", ulps=3, atol=1e-3)


def case_layernorm_pool(make):
    d = make(13)
    n, H, W, Cc = 1, 6, 9, 512
    x = d.tensor(n, H, W, Cc, 2.0)
    g, be = (1 + d.rng.standard_normal(Cc) * 0.1).astype(np.float32), (d.rng.standard_normal(Cc) * 0.1).astype(np.float32)
    out = d.tensor(n, H, W, Cc, fill=False)
    d.call("layernorm", x, d.param(g), d.param(be), out)
    out.check("layernorm", ulps=3, atol=1e-3)

    Cc = 16
    x = d.tensor(2, 8, 12, Cc)
    w = (d.rng.standard_normal((Cc, 16)) * 0.3).astype(np.float32)
    bias = d.rng.standard_normal(Cc).astype(np.float32)
    out = d.tensor(2, 2, 3, Cc, fill=False)
    d.call("pool4", x, d.param(w), d.param(bias), out)
    out.check("pool4", ulps=3, atol=1e-3)


def case_window_attention(make, heads=1, Hn=10, Wn=18, ph=2, pw=2, n_valid=23):
    d = make(14)
    T, Cc = 3, 128 * heads
    q, k, v = (d.tensor(T, Hn, Wn, Cc, 1.5) for _ in range(3))
    kp, vp = d.tensor(T, ph, pw, Cc, 1.5), d.tensor(T, ph, pw, Cc, 1.5)
    valid = d.ints(np.sort(d.rng.choice(180, n_valid, replace=False)))
    n_win = (Hn // 5) * (Wn // 9)
    tind, masked = d.ints([0, 2]), d.ints([(i * 7 + 1) % 3 != 0 for i in range(n_win)])
    out = d.tensor(T, Hn, Wn, Cc, fill=False)
    d.call("window_attention", q, k, v, kp, vp, valid, n_valid, tind, 2, masked, out)
    out.check("window_attention", ulps=4, atol=2e-3)


def case_pred_to_rgb8(make):
    d = make(15)
    x = d.tensor(2, 5, 6, 64, 1.5)
    want, got = d.call("pred_to_rgb8", x)
    assert got.shape == want.shape == (2, 5, 6, 3)
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1 and (got == want).mean() > 0.98


def case_entry_points_reject_bad_arguments(make):
    """the argument checks of the shipped entry points run too: a pitch that is not a multiple of 8, a null tensor"""
    from vsr_b200 import _capi

    d = make(16)
    x, y = d.tensor(1, 4, 4, 12), d.tensor(1, 4, 4, 12)
    with pytest.raises(_capi.VsrError, match="bad arguments"):
        d.real.instnorm(x.real, y.real, 0)
    with pytest.raises(_capi.VsrError, match="bad arguments"):
        d.real.write_extra(x.real, y.real, 10, 5)                                  # 10 + 5 channels do not fit a pitch of 12


CASES = [
    ("frames_and_states", case_frames_and_states, ()),
    ("instnorm", case_instnorm, (0,)), ("instnorm_relu", case_instnorm, (1,)), ("instnorm_128ch", case_instnorm, (1, 128)),
    ("context_split_and_gru", case_context_split_and_gru, ()),
    ("corr_pool_and_lookup", case_corr_pool_and_lookup, ()), ("corr_pool_and_lookup_odd_map", case_corr_pool_and_lookup, (9, 17)),
    ("flow_update_refresh", case_flow_update, (0,)), ("flow_update_add", case_flow_update, (1,)),
    ("convex_upsample", case_convex_upsample, ()),
    ("img_prop_step", case_img_prop_step, ()), ("img_prop_step_wider_than_a_block", case_img_prop_step, (5, 300)),
    ("rfc_input_combine", case_rfc_input_combine, (False,)), ("rfc_input_combine_reversed", case_rfc_input_combine, (True,)),
    ("pad_leaky_taps_extra", case_pad_leaky_taps_extra, ()),
    ("deform_cols", case_deform_cols, (False, False)), ("deform_cols_two_inputs_flow", case_deform_cols, (True, True)),
    ("deform_cols_4_channel_groups", case_deform_cols, (True, True, 8)), ("deform_cols_128ch_16_groups", case_deform_cols, (True, True, 16, 128)),
    ("gen_input_flow_down_masks", case_gen_input_flow_down_masks, ()),
    ("featprop_cond", case_featprop_cond, ()),
    ("unfold_fold", case_unfold_fold, (False,)), ("unfold_fold_gelu", case_unfold_fold, (True,)), ("unfold_fold_11x15", case_unfold_fold, (False, 11, 15)),
    ("unfold_fold_pipeline_map", case_unfold_fold, (False, 32, 48)),
    ("layernorm_pool", case_layernorm_pool, ()),
    ("window_attention", case_window_attention, ()), ("window_attention_two_heads", case_window_attention, (2,)),
    ("window_attention_pipeline_geometry", case_window_attention, (1, 15, 27, 3, 4, 148)),
    ("pred_to_rgb8", case_pred_to_rgb8, ()),
    ("entry_points_reject_bad_arguments", case_entry_points_reject_bad_arguments, ()),
]
