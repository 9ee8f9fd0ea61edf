"""The ProPainter kernels of csrc/pp_ops.cuh have not run on a B200 yet.  Until they do, this file runs their REAL SOURCE on the CPU:
tests/emu/cuda_emu.h defines the handful of CUDA constructs the file uses (__half with round-to-nearest-even, uint4, blockIdx/threadIdx, warp
shuffles and __syncthreads through one OS thread per CUDA thread), tests/emu/pp_emu.cpp launches every kernel with the grid the C ABI uses, and
each result is compared with the numpy stand-in of the runtime (tests/fake_rt.py) that the CPU parity tests of the pipeline are built on.
That closes the gap "the stand-in is a transcription, not the kernel": index arithmetic, layouts and reductions of the kernels themselves are
executed here.  What it cannot show: anything about the device (memory model, launch limits, performance)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from fake_rt import FakeRuntime

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")


@pytest.fixture(scope="module")
def emu():
    out = os.path.join(EMU, "build", "libpp_emu.so")
    srcs = [os.path.join(EMU, "pp_emu.cpp"), os.path.join(EMU, "cuda_emu.h"), os.path.join(HERE, "..", "video-subtitle-remover_b200", "csrc", "pp_ops.cuh")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(s) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-I", os.path.join(EMU, "stubs"), "-I", EMU, srcs[0], "-o", out], check=True)
    return ctypes.CDLL(out)


class Tn:
    def __init__(self, ptr, n, h, w, cp, c=None):
        self.ptr, self.n, self.h, self.w, self.cp, self.c = ptr, n, h, w, cp, cp if c is None else c

    @property
    def pixels(self):
        return self.n * self.h * self.w


class Bench:
    """fake runtime + helpers to mirror its buffers into the flat arrays the emulated kernels take"""

    def __init__(self, seed):
        self.rt = FakeRuntime()
        self.rng = np.random.default_rng(seed)

    def tensor(self, n, h, w, cp, scale=1.0, fill=True):
        t = Tn(self.rt.alloc(n * h * w * cp * 2), n, h, w, cp)
        if fill:
            self.rt._v4(t)[:] = (self.rng.standard_normal((n, h, w, cp)) * scale).astype(np.float16).astype(np.float32)
        return t

    def half(self, t):
        """the tensor's device image: uint16 bits of the fp16 values, [n,h,w,cp] contiguous"""
        return np.ascontiguousarray(self.rt._v4(t).astype(np.float16)).view(np.uint16).copy()

    def f32(self, arr):
        arr = np.ascontiguousarray(arr, np.float32)
        return self.rt.upload_bytes(arr), arr.copy()

    def u8(self, arr):
        arr = np.ascontiguousarray(arr, np.uint8)
        return self.rt.upload_bytes(arr), arr.copy()

    def ints(self, arr):
        arr = np.ascontiguousarray(arr, np.int32)
        return self.rt.upload_ints(arr), arr.copy()


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def Z(n):
    return ctypes.c_size_t(int(n))


def as_f32(bits):
    return bits.view(np.float16).astype(np.float32)


def check(got_bits, fake_view, what, ulps=2, atol=0.0):
    got = as_f32(got_bits).reshape(fake_view.shape)
    want = fake_view.astype(np.float16).astype(np.float32)
    tol = ulps * np.maximum(np.abs(want), 2.0 ** -14) * 2.0 ** -10 + atol      # atol: sums with cancellation (fp32 summation order differs)
    bad = np.abs(got - want) > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.size} differ, worst {np.abs(got - want).max():.4g}"


def test_frames_and_states(emu):
    b = Bench(0)
    T, H, W = 2, 6, 10
    frames = [b.rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(T)]
    y = b.tensor(T, H, W, 8, fill=False)
    b.rt.frames(frames, y)
    out = np.zeros((T, H, W, 8), np.uint16)
    emu.emu_frames_to_half(P(np.ascontiguousarray(np.stack(frames))), Z(T * H * W), P(out))
    check(out, b.rt._v4(y), "frames", ulps=0)

    mask_h, mask = b.u8((b.rng.random((H, W)) > 0.6) * 255)
    st = b.tensor(T, H, W, 8, fill=False)
    b.rt.prop_state(y, mask_h, None, st)
    got = np.zeros((T, H, W, 8), np.uint16)
    emu.emu_state_init(P(b.half(y)), P(mask), Z(H * W), Z(T * H * W), P(got))
    check(got, b.rt._v4(st), "state_init", ulps=0)

    prop = b.tensor(T, H, W, 8)
    st2 = b.tensor(T, H, W, 8, fill=False)
    prop_bits = b.half(prop)
    b.rt.prop_state(y, mask_h, prop, st2)
    got = np.zeros((T, H, W, 8), np.uint16)
    emu.emu_state_compose(P(b.half(y)), P(mask), P(prop_bits), Z(H * W), Z(T * H * W), P(got))
    check(got, b.rt._v4(st2), "state_compose", ulps=0)


@pytest.mark.parametrize("relu", [0, 1])
def test_instnorm(emu, relu):
    b = Bench(1)
    N, H, W, cp = 2, 9, 13, 64
    x = b.tensor(N, H, W, cp, scale=3.0)
    b.rt._v4(x)[:] += np.linspace(-2, 2, cp, dtype=np.float32)
    b.rt._v4(x)[:] = b.rt._v4(x).astype(np.float16).astype(np.float32)
    y = b.tensor(N, H, W, cp, fill=False)
    xb = b.half(x)
    b.rt.instnorm(x, y, relu)
    got = np.zeros_like(xb)
    mean, rstd = np.zeros(N * cp, np.float32), np.zeros(N * cp, np.float32)
    emu.emu_instnorm(P(xb), N, Z(H * W), cp, relu, P(mean), P(rstd), P(got))
    v = as_f32(xb).reshape(N, H * W, cp)
    np.testing.assert_allclose(mean.reshape(N, cp), v.mean(1), atol=1e-4)
    np.testing.assert_allclose(rstd.reshape(N, cp), 1 / np.sqrt(v.var(1) + 1e-5), rtol=1e-4)
    check(got, b.rt._v4(y), "instnorm", ulps=3)


def test_context_split_and_gru(emu):
    b = Bench(2)
    N, H, W = 1, 5, 7
    px = N * H * W
    x = b.tensor(N, H, W, 256)
    net, inp = b.tensor(N, H, W, 384), b.tensor(N, H, W, 384)
    nb, ib = b.half(net), b.half(inp)
    b.rt.context_split(x, net, inp)
    emu.emu_context_split(P(b.half(x)), Z(px), P(nb), 384, P(ib), 384)
    check(nb, b.rt._v4(net), "context_split.net")
    check(ib, b.rt._v4(inp), "context_split.inp")

    r, hsrc, out = b.tensor(N, H, W, 128, 2.0), b.tensor(N, H, W, 384), b.tensor(N, H, W, 384)
    ob, rb, hb = b.half(out), b.half(r), b.half(hsrc)
    b.rt.gru_rh(r, hsrc, out)
    emu.emu_gru_rh(P(rb), 128, P(hb), 384, P(ob), 384, Z(px))
    check(ob, b.rt._v4(out), "gru_rh")

    z, q, hio = b.tensor(N, H, W, 128, 2.0), b.tensor(N, H, W, 128, 2.0), b.tensor(N, H, W, 384)
    zb, qb, hb = b.half(z), b.half(q), b.half(hio)
    b.rt.gru_update(z, q, hio)
    emu.emu_gru_update(P(zb), 128, P(qb), 128, P(hb), 384, Z(px))
    check(hb, b.rt._v4(hio), "gru_update")


def test_corr_pool_and_lookup(emu):
    b = Bench(3)
    hh, ww = 8, 12                      # 1/8-resolution map; level l has (hh >> l, ww >> l) targets per source pixel
    px = hh * ww
    levels, arrays = [], []
    H, W = hh, ww
    pitch = (H * W + 7) // 8 * 8
    src = Tn(b.rt.alloc(px * pitch * 2), 1, 1, px, pitch)
    b.rt._v4(src)[:] = b.rng.standard_normal((1, 1, px, pitch)).astype(np.float16).astype(np.float32)
    levels.append((src.ptr, H, W, pitch))
    arrays.append(b.half(src))
    for _ in range(3):
        oh, ow = H // 2, W // 2
        op = (oh * ow + 7) // 8 * 8
        dst = Tn(b.rt.alloc(px * op * 2), 1, 1, px, op)
        got = np.zeros((px, op), np.uint16)
        b.rt.corr_pool(levels[-1][0], px, H, W, pitch, dst.ptr, op)
        emu.emu_corr_pool(P(arrays[-1]), Z(px), H, W, pitch, P(got), op)
        check(got.reshape(1, 1, px, op)[..., : oh * ow], b.rt._v4(dst)[..., : oh * ow], f"corr_pool {H}x{W}")
        b.rt._v4(dst)[:] = b.rt._v4(dst).astype(np.float16).astype(np.float32)
        levels.append((dst.ptr, oh, ow, op))
        arrays.append(b.half(dst))
        H, W, pitch = oh, ow, op
    flow_h, flow = b.f32(b.rng.standard_normal((px, 2)) * 3)
    out = b.tensor(1, hh, ww, 384, fill=False)
    b.rt.corr_lookup(levels, flow_h, hh, ww, px, out)
    got = np.zeros((1, hh, ww, 384), np.uint16)
    hs, ws, ps = (np.array([lv[i] for lv in levels], np.int32) for i in (1, 2, 3))
    emu.emu_corr_lookup(P(arrays[0]), P(arrays[1]), P(arrays[2]), P(arrays[3]), P(hs), P(ws), P(ps), P(flow), hh, ww, Z(px), P(got), 384)
    check(got[..., :324], b.rt._v4(out)[..., :324], "corr_lookup", ulps=3)
    assert np.abs(as_f32(got[..., :324])).max() > 0.5


@pytest.mark.parametrize("add", [0, 1])
def test_flow_update(emu, add):
    b = Bench(4)
    N, H, W = 2, 4, 6
    px = N * H * W
    f_h, f = b.f32(b.rng.standard_normal((px, 2)) * 4)
    delta = b.tensor(N, H, W, 64)
    f16, a, c = b.tensor(N, H, W, 8), b.tensor(N, H, W, 384), b.tensor(N, H, W, 384)
    fb, ab, cb, db = b.half(f16), b.half(a), b.half(c), b.half(delta)
    b.rt.flow_update(f_h, delta, f16, a, c, 382, add)
    emu.emu_flow_update(P(f), P(db), 64, P(fb), P(ab), P(cb), 384, 382, Z(px), add)
    np.testing.assert_array_equal(f, b.rt._raw32(f_h, px * 2).reshape(px, 2))
    for bits, t, name in ((fb, f16, "flow16"), (ab, a, "dst_a"), (cb, c, "dst_b")):
        check(bits, b.rt._v4(t), "flow_update." + name, ulps=0)


def test_convex_upsample(emu):
    b = Bench(5)
    N, h, w = 2, 4, 5
    f_h, f = b.f32(b.rng.standard_normal((N * h * w, 2)) * 3)
    mask = b.tensor(N, h, w, 576, 2.0)
    out_h, _ = b.f32(np.zeros(N * 2 * 64 * h * w))
    b.rt.convex_upsample(f_h, mask, N, h, w, out_h)
    got = np.zeros(N * 2 * 64 * h * w, np.float32)
    emu.emu_convex_upsample(P(f), P(b.half(mask)), 576, N, h, w, P(got))
    np.testing.assert_allclose(got, b.rt._raw32(out_h, got.size), rtol=2e-5, atol=2e-5)


def test_img_prop_step(emu):
    b = Bench(6)
    H, W = 12, 20
    prev, cur = b.tensor(1, H, W, 8), b.tensor(1, H, W, 8)
    b.rt._v4(prev)[..., 3] = b.rng.random((1, H, W)) > 0.93       # few holes left in the source frame, half of the current frame missing
    b.rt._v4(cur)[..., 3] = b.rng.random((1, H, W)) > 0.5
    fp_h, fp = b.f32(b.rng.standard_normal((2, H, W)) * 2)
    back = -fp + b.rng.standard_normal((2, H, W)).astype(np.float32) * 0.4        # consistent for some pixels, not for others
    fc_h, fc = b.f32(back)
    out = b.tensor(1, H, W, 8, fill=False)
    b.rt.img_prop_step(prev, cur, fp_h, fc_h, out)
    got = np.zeros((1, H, W, 8), np.uint16)
    emu.emu_img_prop_step(P(b.half(prev)), P(b.half(cur)), P(fp), P(fc), H, W, P(got))
    check(got, b.rt._v4(out), "img_prop_step", ulps=0)
    changed = (b.rt._v4(out)[..., 3] != b.rt._v4(cur)[..., 3]).mean()
    assert 0.02 < changed < 0.98                                                  # both branches of the fill decision were taken


@pytest.mark.parametrize("reverse", [0, 1])
def test_rfc_input_combine(emu, reverse):
    b = Bench(7)
    N, H, W = 3, 5, 6
    fl_h, fl = b.f32(b.rng.standard_normal((N, 2, H, W)) * 3)
    m_h, m = b.u8((b.rng.random((H, W)) > 0.5) * 255)
    out = b.tensor(N, H, W, 8)
    b.rt.rfc_input(fl_h, m_h, N, H, W, reverse, out)
    got = np.zeros((N, H, W, 8), np.uint16)
    emu.emu_rfc_input(P(fl), P(m), N, Z(H * W), reverse, P(got))
    check(got, b.rt._v4(out), "rfc_input", ulps=0)

    pred = b.tensor(N, H, W, 64)
    o_h, _ = b.f32(np.zeros(N * 2 * H * W))
    b.rt.rfc_combine(pred, fl_h, m_h, N, H, W, reverse, o_h)
    got = np.zeros(N * 2 * H * W, np.float32)
    emu.emu_rfc_combine(P(b.half(pred)), 64, P(fl), P(m), N, Z(H * W), reverse, P(got))
    np.testing.assert_array_equal(got, b.rt._raw32(o_h, got.size))


def test_pad_leaky_taps_extra(emu):
    b = Bench(8)
    T, H, W, cp = 3, 5, 7, 16
    x = b.tensor(T, H, W, cp)
    y = b.tensor(T, H + 3, W + 4, cp, fill=False)
    b.rt.pad_replicate(x, y, 1, 2)
    got = np.zeros((T, H + 3, W + 4, cp), np.uint16)
    emu.emu_pad_replicate(P(b.half(x)), T, H, W, cp, P(got), H + 3, W + 4, 1, 2)
    check(got, b.rt._v4(y), "pad_replicate", ulps=0)

    xb = b.half(x)
    b.rt.leaky(x, 0.2)
    emu.emu_leaky(P(xb), Z(xb.size // 8), ctypes.c_float(0.2))
    check(xb, b.rt._v4(x), "leaky", ulps=1)

    taps = b.tensor(T, H, W, 64)
    tb = b.half(taps)
    xb = b.half(x)
    b.rt.temporal_taps(x, taps)
    emu.emu_temporal_taps(P(xb), T, Z(H * W), cp, P(tb), 64)
    check(tb[..., : 3 * cp], b.rt._v4(taps)[..., : 3 * cp], "temporal_taps", ulps=0)

    dst = b.tensor(T, H, W, 64)
    db = b.half(dst)
    src8 = b.tensor(T, H, W, 8)                                  # the kernel's contract: an 8-channel source
    b.rt.write_extra(src8, dst, 37, 5)
    emu.emu_write_extra(P(b.half(src8)), P(db), 64, 37, 5, Z(T * H * W))
    check(db, b.rt._v4(dst), "write_extra", ulps=0)


@pytest.mark.parametrize("two_inputs,with_flow", [(False, False), (True, True)])
def test_deform_cols(emu, two_inputs, with_flow):
    b = Bench(9)
    n, H, W, C, G = 2, 6, 7, 32, 4
    ca = 16 if two_inputs else C
    xa = b.tensor(n, H, W, 64)
    xb = b.tensor(n, H, W, 64) if two_inputs else None
    om = b.tensor(n, H, W, 128, 0.7)
    cols = b.tensor(n, H, W, 9 * C + 32)
    cb = b.half(cols)
    fl_h, fl = b.f32(b.rng.standard_normal((n * H * W, 2)) * 2) if with_flow else (0, None)
    b.rt.deform_cols(xa, ca, xb, C, G, om, 3.0, fl_h, cols)
    emu.emu_deform_cols(P(b.half(xa)), 64, ca, P(b.half(xb)) if two_inputs else None, 64, C, G, P(b.half(om)), 128, ctypes.c_float(3.0), P(fl), H, W,
                        Z(n * H * W), P(cb), 9 * C + 32)
    check(cb, b.rt._v4(cols), "deform_cols", ulps=4)


def test_gen_input_flow_down_masks(emu):
    b = Bench(10)
    T, H, W = 4, 8, 12
    state = b.tensor(T, H, W, 8)
    m_h, m = b.u8((b.rng.random((H, W)) > 0.5) * 255)
    ids_h, ids = b.ints([2, 0, 3])
    gin = b.tensor(3, H, W, 8)
    b.rt.gen_input(state, m_h, ids_h, 3, gin)
    got = np.zeros((3, H, W, 8), np.uint16)
    emu.emu_gen_input(P(b.half(state)), P(m), P(ids), 3, Z(H * W), P(got))
    check(got, b.rt._v4(gin), "gen_input", ulps=0)

    fl_h, fl = b.f32(b.rng.standard_normal((T, 2, H, W)) * 3)
    o_h, _ = b.f32(np.zeros(3 * (H // 4) * (W // 4) * 2))
    b.rt.flow_down4(fl_h, ids_h, 3, H, W, o_h)
    gotf = np.zeros(3 * (H // 4) * (W // 4) * 2, np.float32)
    emu.emu_flow_down4(P(fl), P(ids), 3, H, W, P(gotf))
    np.testing.assert_allclose(gotf, b.rt._raw32(o_h, gotf.size), rtol=1e-6, atol=1e-6)

    pm = b.tensor(3, H // 4, W // 4, 8)
    b.rt.prop_masks(gin, pm)
    got = np.zeros((3, H // 4, W // 4, 8), np.uint16)
    emu.emu_prop_masks(P(b.half(gin)), 3, H, W, P(got))
    check(got, b.rt._v4(pm), "prop_masks", ulps=0)


def test_featprop_cond(emu):
    b = Bench(11)
    H, W, C = 7, 9, 16
    prop, cur, masks = b.tensor(1, H, W, C), b.tensor(1, H, W, C), b.tensor(1, H, W, 8)
    fp_h, fp = b.f32(b.rng.standard_normal((H * W, 2)) * 2)
    fc_h, fc = b.f32(-fp + b.rng.standard_normal((H * W, 2)).astype(np.float32) * 0.4)
    cond = b.tensor(1, H, W, 64)
    cb = b.half(cond)
    b.rt.featprop_cond(prop, cur, fp_h, fc_h, masks, cond)
    emu.emu_featprop_cond(P(b.half(prop)), P(b.half(cur)), C, P(fp), P(fc), P(b.half(masks)), H, W, P(cb), 64)
    check(cb[..., : 2 * C + 5], b.rt._v4(cond)[..., : 2 * C + 5], "featprop_cond", ulps=3)
    valid = b.rt._v4(cond)[..., 2 * C + 2]
    assert 0.05 < valid.mean() < 0.95


@pytest.mark.parametrize("gelu", [0, 1])
def test_unfold_fold(emu, gelu):
    b = Bench(12)
    n, h, w, C = 2, 10, 13, 8
    fh, fw = (h + 6 - 7) // 3 + 1, (w + 6 - 7) // 3 + 1
    x = b.tensor(n, h, w, C)
    tok = b.tensor(n, fh, fw, 49 * C + 8)
    tb = b.half(tok)
    b.rt.unfold7s3(x, tok, bool(gelu))
    emu.emu_unfold7s3(P(b.half(x)), n, h, w, C, P(tb), 49 * C + 8, gelu)
    check(tb[..., : 49 * C], b.rt._v4(tok)[..., : 49 * C], "unfold7s3", ulps=2)

    b.rt._v4(tok)[:] = b.rt._v4(tok).astype(np.float16).astype(np.float32)
    for norm in (0, 1):
        out = b.tensor(n, h, w, C, fill=False)
        b.rt.fold7s3(tok, out, C, norm)
        got = np.zeros((n, h, w, C), np.uint16)
        emu.emu_fold7s3(P(b.half(tok)), n, h, w, C, 49 * C + 8, norm, P(got))
        check(got, b.rt._v4(out), f"fold7s3 normalise={norm}", ulps=3)


def test_layernorm_pool(emu):
    b = Bench(13)
    n, H, W, C = 1, 6, 9, 512
    x = b.tensor(n, H, W, C, 2.0)
    g_h, g = b.f32(1 + b.rng.standard_normal(C) * 0.1)
    be_h, be = b.f32(b.rng.standard_normal(C) * 0.1)
    # the stand-in reads its parameters as contiguous fp32 vectors
    g_h, be_h = b.rt.upload_f32(g), b.rt.upload_f32(be)
    out = b.tensor(n, H, W, C, fill=False)
    b.rt.layernorm(x, g_h, be_h, out)
    got = np.zeros((n, H, W, C), np.uint16)
    emu.emu_layernorm(P(b.half(x)), Z(n * H * W), C, P(g), P(be), P(got))
    check(got, b.rt._v4(out), "layernorm", ulps=3)

    C = 16
    x = b.tensor(2, 8, 12, C)
    w = b.rng.standard_normal((C, 16)).astype(np.float32) * 0.3
    bias = b.rng.standard_normal(C).astype(np.float32)
    out = b.tensor(2, 2, 3, C, fill=False)
    b.rt.pool4(x, b.rt.upload_f32(w), b.rt.upload_f32(bias), out)
    got = np.zeros((2, 2, 3, C), np.uint16)
    emu.emu_pool4(P(b.half(x)), 2, 8, 12, C, P(w), P(bias), P(got))
    check(got, b.rt._v4(out), "pool4", ulps=3)


def test_window_attention(emu):
    b = Bench(14)
    T, Hn, Wn, C, ph, pw = 3, 10, 18, 128, 2, 2
    q, k, v = (b.tensor(T, Hn, Wn, C, 1.5) for _ in range(3))
    kp, vp = b.tensor(T, ph, pw, C, 1.5), b.tensor(T, ph, pw, C, 1.5)
    valid = np.sort(b.rng.choice(180, 23, replace=False)).astype(np.int32)
    tind = np.array([0, 2], np.int32)
    masked = np.array([1, 0, 0, 1], np.int32)
    out = b.tensor(T, Hn, Wn, C, fill=False)
    b.rt.window_attention(q, k, v, kp, vp, b.rt.upload_ints(valid), valid.size, b.rt.upload_ints(tind), tind.size, b.rt.upload_ints(masked), out)
    got = np.zeros((T, Hn, Wn, C), np.uint16)
    emu.emu_window_attention(P(b.half(q)), P(b.half(k)), P(b.half(v)), P(b.half(kp)), P(b.half(vp)), T, Hn, Wn, C, ph, pw, P(valid), valid.size, P(tind), tind.size,
                             P(masked), P(got))
    check(got, b.rt._v4(out), "window_attention", ulps=4, atol=2e-3)


def test_pred_to_rgb8(emu):
    b = Bench(15)
    x = b.tensor(2, 5, 6, 64, 1.5)
    want = b.rt.pred_to_rgb8(x)
    got = np.zeros((2, 5, 6, 3), np.uint8)
    emu.emu_pred_to_rgb8(P(b.half(x)), 64, Z(60), P(got))
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1 and (got == want).mean() > 0.98
