"""GPU: the LAMA path — the LAMA-only runtime operators against numpy, and `LamaInpaint` against the
oracle (oracle/lama_oracle.py, pinned bit-identical to the reference's TorchScript module on the CPU).
Tolerance (stated): the device network multiplies fp16 operands (fp32 accumulate, fp32 FFT, fp32 master copy of the residual
stream) through 18 residual FFC blocks.  Inside the hole, against the oracle's u8 output: PSNR >= 45 dB and |diff| <= 10 of
255 with the reference's weights (measured 52-57 dB, max 2-5; profiles/lama_r1.json), PSNR >= 38 dB and |diff| <= 12 with the
seeded random weights (un-trained weights amplify rounding more).  Outside the hole the output is the input image through
the reference's own fp32 blend: bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import lama_oracle as L
from oracle import sttn_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rt(capi):
    if capi.lib().vsr_device_count() < 1:
        pytest.fail("GPU tests need a B200 (sm_100) device")
    from vsr_b200.lama_inpaint import _LamaRuntime

    r = _LamaRuntime("cuda:0")
    yield r
    r.close()


def _tensor(rt, arr, cp=None):
    """numpy [h,w,c] -> device NHWC fp16 tensor (pitch cp)."""
    from vsr_b200.dbnet import _Tensor, _r

    h, w, c = arr.shape
    cp = cp or _r(c, 64)
    host = np.zeros((h, w, cp), np.float16)
    host[:, :, :c] = arr
    t = _Tensor(rt.alloc(host.nbytes), c, h, w, cp)
    rt.L.vsr_rt_upload(rt.h, t.ptr, host.ctypes.data_as(C.c_void_p), host.nbytes)
    return t


def _empty(rt, c, h, w, cp=None):
    from vsr_b200.dbnet import _Tensor, _r

    cp = cp or _r(c, 64)
    return _Tensor(rt.alloc(h * w * cp * 2), c, h, w, cp)


def test_pad_upsample_add_slices(rt):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((9, 13, 64)).astype(np.float16)
    tx = _tensor(rt, x)
    ty = _empty(rt, 64, 9 + 4, 13 + 6)
    rt.pad(tx, ty, 1, 2)
    assert np.array_equal(rt.download(ty), np.pad(x, ((1, 3), (2, 4), (0, 0)), mode="reflect"))
    rt.pad(tx, ty, 1, 2, 0)
    assert np.array_equal(rt.download(ty), np.pad(x, ((1, 3), (2, 4), (0, 0))))
    tz = _empty(rt, 64, 18, 26)
    rt.zero_upsample(tx, tz)
    z = np.zeros((18, 26, 64), np.float16)
    z[::2, ::2] = x
    assert np.array_equal(rt.download(tz), z)
    from vsr_b200.lama_inpaint import _view

    a = rng.standard_normal((9, 13, 128)).astype(np.float16)
    b = rng.standard_normal((9, 13, 64)).astype(np.float16)
    ta, tb, to = _tensor(rt, a), _tensor(rt, b), _tensor(rt, np.full((9, 13, 192), 7, np.float16))
    rt.add_slices(1, _view(ta, 64, 64), tb, _view(to, 128, 64), 64)
    want = np.full((9, 13, 192), 7, np.float32)
    want[:, :, 128:] = np.maximum(a[:, :, 64:].astype(np.float32) + b.astype(np.float32), 0)
    assert np.array_equal(rt.download(to), want.astype(np.float16))
    assert not rt.overflow()


def test_fourier_unit_layout_and_roundtrip(rt):
    rng = np.random.default_rng(1)
    h, w, c = 45, 240, 192     # the /8 grid of a 1080p LAMA strip: non power-of-two sizes
    x = rng.standard_normal((h, w, c)).astype(np.float16)
    tx, tf, ty = _tensor(rt, x), _empty(rt, 2 * c, h, w // 2 + 1), _empty(rt, c, h, w)
    rt.fft_r2c(tx, tf)
    f = np.fft.rfft2(x.astype(np.float64), axes=(0, 1), norm="ortho")
    want = np.stack([f.real, f.imag], -1).reshape(h, w // 2 + 1, 2 * c)
    got = rt.download(tf).astype(np.float64)
    assert np.abs(got - want).max() <= 2e-3 * np.abs(want).max()          # fp16 storage of an fp32 transform
    rt.fft_c2r(tf, ty)
    back = rt.download(ty).astype(np.float64)
    assert np.abs(back - x).max() <= 6e-3
    assert not rt.overflow()


def test_conv_on_padded_grid_with_cropped_store(rt):
    """reflect-padded 3x3 conv and the stride-2 variant as the LAMA graph issues them, against torch."""
    import torch
    import torch.nn.functional as F
    from vsr_b200.lama_inpaint import _view

    rng = np.random.default_rng(2)
    h, w, cin, cout = 12, 20, 128, 64
    x = (rng.standard_normal((h, w, cin)) * 0.5).astype(np.float16)
    wt = (rng.standard_normal((cout, 64, 3, 3)) / 24).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    xt = torch.from_numpy(x.astype(np.float32)).permute(2, 0, 1)[None]
    wq = torch.from_numpy(wt).half().float()
    tx, tp = _tensor(rt, x), _empty(rt, cin, h + 2, w + 2)
    rt.pad(tx, tp, 1, 1)
    # (a) channels 64..127 only (slice view), result into channels 64.. of a 128-channel output
    lid = rt.conv_create(wt, bias, cout, 64, tp.cp, 3, 3, 1, 1, 1, 1, 1, False)
    ty = _tensor(rt, np.full((h, w, 128), 3, np.float16))
    rt.conv_ex(lid, _view(tp, 64, 64), ty, 1, 64, (1, 1))
    want = F.relu(F.conv2d(F.pad(xt[:, 64:], (1, 1, 1, 1), mode="reflect"), wq, torch.from_numpy(bias)))[0].permute(1, 2, 0).numpy()
    got = rt.download(ty).astype(np.float32)
    assert np.array_equal(got[:, :, :64], np.full((h, w, 64), 3, np.float32))
    assert np.abs(got[:, :, 64:] - want).max() <= 2e-2
    # (b) stride 2 through the shifted padding P''[a] = x[reflect(a - 2)]
    w2 = (rng.standard_normal((cout, cin, 3, 3)) / 34).astype(np.float32)
    tq, tz = _empty(rt, cin, h + 4, w + 4), _empty(rt, cout, h // 2, w // 2)
    rt.pad(tx, tq, 2, 2)
    lid2 = rt.conv_create(w2, bias, cout, cin, tq.cp, 3, 3, 2, 1, 1, 1, 1, False)
    rt.conv_ex(lid2, tq, tz, 0, 0, (1, 1))
    want2 = F.conv2d(F.pad(xt, (1, 1, 1, 1), mode="reflect"), torch.from_numpy(w2).half().float(), torch.from_numpy(bias), stride=2)[0]
    assert np.abs(rt.download(tz).astype(np.float32) - want2.permute(1, 2, 0).numpy()).max() <= 2e-2
    assert not rt.overflow()


def _check(got, want, img, mask, min_psnr=38.0, max_diff=12):
    hole = mask > 0
    assert np.array_equal(got[~hole], want[~hole])                    # the reference's own blend, bit-exact
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    psnr = O.psnr_u8(got[hole].astype(np.float32), want[hole].astype(np.float32))
    assert psnr >= min_psnr and d.max() <= max_diff, (psnr, int(d.max()))
    assert np.abs(got[hole].astype(np.int32) - img[hole]).mean() > 5   # the hole really was repainted


@pytest.fixture(scope="module")
def lama_rand(capi):
    from vsr_b200.lama_inpaint import LamaInpaint

    w = L.random_weights(3)
    return LamaInpaint("cuda:0", {k: v.numpy() for k, v in w.items()}), w


def test_inpaint_vs_oracle_random_weights(lama_rand):
    eng, w = lama_rand
    rng = np.random.default_rng(7)
    for hw in ((70, 100), (128, 192)):
        img = rng.integers(0, 256, hw + (3,), dtype=np.uint8)
        mask = np.zeros(hw, np.uint8)
        mask[hw[0] // 3: 2 * hw[0] // 3, 10: hw[1] - 8] = 255
        got = eng.inpaint(img, mask)
        _check(got, L.inpaint(w, img, mask), img, mask)
        assert np.array_equal(eng.inpaint(img, mask), got)             # deterministic, graph replay


def test_call_strips_vs_oracle_random_weights(lama_rand):
    eng, w = lama_rand
    H, W, T = 270, 480, 6          # strips go through the network 4 + 2 per launch
    frames = O.synthetic_clip(T, H, W, seed=31)
    keep = [f.copy() for f in frames]
    mask = O.default_mask(H, W)
    out = eng(frames, mask)
    assert all(np.array_equal(a, b) for a, b in zip(frames, keep))
    want = L.lama_call(w, frames, mask)
    (y0, y1, _, _), = O.get_inpaint_area_by_mask(W, H, int(W * 3 / 16), mask)
    for o, r, f in zip(out, want, keep):
        assert np.array_equal(o[:y0], f[:y0]) and np.array_equal(o[y1:], f[y1:])
        _check(o[y0:y1], r[y0:y1], f[y0:y1], mask[y0:y1])
    assert sorted(k[0] for k in eng.model._programs if k[1:] == (96, 480)) == [2, 4]
    single = eng.inpaint(keep[5][y0:y1], mask[y0:y1])                 # batching does not change a frame's result
    assert np.abs(single.astype(np.int32) - out[5][y0:y1]).max() <= 1


def test_golden_real_weights(capi):
    """The unmodified reference's outputs (tests/golden/lama_real.npz) with the reference's weights, when they are staged."""
    import sys

    path = next((p for p in (os.path.join(ROOT, "weights", "big-lama", "big-lama.pt"), os.path.join(ROOT, "weights", "big-lama", "big-lama.npz"))
                 if os.path.exists(p)), None)
    if path is None:
        pytest.skip("big-lama weights not staged under weights/big-lama (206 MB; tools/stage_weights.py --lama)")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_golden_lama import inputs
    from vsr_b200.lama_inpaint import LamaInpaint

    eng = LamaInpaint("cuda:0", path)
    z = np.load(os.path.join(GOLDEN, "lama_real.npz"))
    img, m, frames, mask = inputs()
    _check(eng.inpaint(img, m), z["single"], img, m, 45.0, 10)
    out = eng(frames, mask)
    for o, r, f in zip(out, z["call"], frames):
        hole_rows = np.flatnonzero(mask.any(1))
        y0, y1 = hole_rows[0], hole_rows[-1] + 1
        assert np.array_equal(o[:y0 - 40], f[:y0 - 40])              # rows far from the strip are the caller's pixels
        # a 48-row strip that is 40 % hole is ill-conditioned: rounding ONLY the conv kernels to fp16 in the fp32 oracle already
        # moves it to 44.3 dB / max 8 (the 70x100 image: 57.4 dB / max 1) — measured on the CPU, DESIGN.md §1.2
        _check(o[y0:y1], r[y0:y1], f[y0:y1], mask[y0:y1], 38.0, 16)


def test_config1_vs_reference_golden(capi):
    """BASELINE config 1: `LamaInpaint.inpaint` on the 512x512 synthetic image (texture seed 0, hole rows 400-470, cols 60-450) and
    on frame 0 of the reference's test/test.mp4 with test/test.png, against the unmodified reference on the CPU (tools/make_golden_configs.py).
    Bar: >= 45 dB and |diff| <= 10 inside the hole, bit-exact outside."""
    path = os.path.join(ROOT, "weights", "big-lama", "big-lama.npz")
    if not os.path.exists(path):
        pytest.skip("big-lama weights not staged under weights/big-lama")
    from vsr_b200.lama_inpaint import LamaInpaint

    eng = LamaInpaint("cuda:0", path)
    z = np.load(os.path.join(GOLDEN, "config1_lama.npz"))
    img = O.synthetic_clip(1, 512, 512, seed=0)[0]
    m = np.zeros((512, 512), np.uint8)
    m[400:470, 60:450] = 255
    _check(eng.inpaint(img, m), z["out512"], img, m, 45.0, 10)
    _check(eng.inpaint(z["test_frame0"], z["test_mask"]), z["test_out"], z["test_frame0"], z["test_mask"], 45.0, 10)
