"""GPU: the STTN-det path through `STTNDetInpaint` against the oracle and the golden
vectors of the unmodified reference.  Same tolerance as tests/test_gpu_sttn.py (PSNR >= 45 dB, max |diff| <= 6)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import sttn_det_oracle as D
from oracle import sttn_oracle as O

pytestmark = pytest.mark.gpu


def _check(got, want):
    got, want = np.stack(got).astype(np.float32), np.stack(want).astype(np.float32)
    assert O.psnr_u8(got, want) >= 45.0, f"psnr {O.psnr_u8(got, want):.2f}"
    assert np.abs(got - want).max() <= 6


@pytest.fixture(scope="module")
def det_rand(capi):
    if capi.lib().vsr_device_count() < 1:
        pytest.fail("GPU tests need a B200 (sm_100) device")
    from vsr_b200 import STTNDetInpaint

    w = O.random_weights(1)
    return STTNDetInpaint("cuda:0", {k: v.numpy() for k, v in w.items()}), w


def test_det_strip_vs_oracle_random_weights(det_rand):
    eng, w = det_rand
    assert (eng.model_input_width, eng.model_input_height) == (432, 240)
    H, W, T = 360, 640, 8
    frames = O.synthetic_clip(T, H, W, seed=17)
    mask = O.default_mask(H, W)
    scaled = [O.cv2_resize_linear_u8(np.ascontiguousarray(f[160:337]), 432, 240) for f in frames]
    msmall = O.cv2_resize_linear_u8(np.ascontiguousarray(mask[160:337]), 432, 240)
    got = eng.inpaint([s.copy() for s in scaled], [msmall] * T)
    want = D.inpaint_strip(w, scaled, [msmall] * T)
    assert [g.dtype for g in got] == [x.dtype for x in want]
    _check(got, want)
    # outside the (dilated) mask the comp is the input frame itself, bit-exact
    keep = msmall == 0
    for g, s in zip(got, scaled):
        assert np.array_equal(g[keep].astype(np.uint8), s[:, :, ::-1][keep])


def test_det_call_vs_oracle_random_weights(det_rand):
    eng, w = det_rand
    H, W, T = 270, 480, 6
    frames = O.synthetic_clip(T, H, W, seed=18)
    keep = [f.copy() for f in frames]
    mask = O.default_mask(H, W)
    out = eng(frames, mask)
    assert all(np.array_equal(a, b) for a, b in zip(frames, keep))
    want = D.det_call(w, frames, mask)
    _check(out, want)
    (y0, y1, _, _), = O.get_inpaint_area_by_mask(W, H, D.split_height(H, W), mask)
    for o, f in zip(out, keep):
        assert np.array_equal(o[:y0], f[:y0]) and np.array_equal(o[y1:], f[y1:])
    # portrait frames use int(H*5/9) (sttn_det_inpaint.py:48-49)
    Hp, Wp = 480, 270
    fp = O.synthetic_clip(3, Hp, Wp, seed=19)
    mp = O.default_mask(Hp, Wp)
    _check(eng(fp, mp), D.det_call(w, fp, mp))


def test_det_vs_reference_golden_real_weights(capi):
    p = os.path.join(ROOT, "weights", "sttn-det", "sttn.pth")
    if not os.path.exists(p):
        pytest.skip("sttn-det checkpoint not staged under weights/")
    from vsr_b200 import STTNDetInpaint

    eng = STTNDetInpaint("cuda:0", p)
    z = np.load(os.path.join(GOLDEN, "sttn_det_real.npz"))
    H, W, T = int(z["H"]), int(z["W"]), int(z["T"])
    frames = O.synthetic_clip(T, H, W, seed=int(z["seed"]))
    mask = O.default_mask(H, W)
    y0, y1 = z["areas"][0][:2]
    scaled = [O.cv2_resize_linear_u8(np.ascontiguousarray(f[y0:y1]), 432, 240) for f in frames]
    got = eng.inpaint(scaled, [z["mask_small"]] * T)
    _check(got, list(z["comps"]))
    out = eng(frames, mask)
    _check([o[y0:y1] for o in out], list(z["strip_out_plain"]))


def test_config4_whole_batch_vs_reference_golden(capi):
    """BASELINE config 4 (inpaint half) at full size: one 46-frame 1080p `batch_generator` batch through the unmodified reference's
    `STTNDetInpaint.__call__` on the CPU (tools/make_golden_configs.py): three stored frames (every second row / column of the 533-row
    strip) pixel by pixel, all 46 through their strip sums."""
    p = os.path.join(ROOT, "weights", "sttn-det", "sttn.pth")
    if not os.path.exists(p):
        pytest.skip("sttn-det checkpoint not staged under weights/")
    from vsr_b200 import STTNDetInpaint

    eng = STTNDetInpaint("cuda:0", p)
    z = np.load(os.path.join(GOLDEN, "config4_sttn_det_1080p.npz"))
    H, W, T = int(z["H"]), int(z["W"]), int(z["T"])
    frames = O.synthetic_clip(T, H, W, seed=int(z["seed"]))
    out = eng(frames, O.default_mask(H, W))
    y0, y1 = (int(v) for v in z["rows"])
    _check([out[int(i)][y0:y1:2, ::2] for i in z["frames"]], list(z["out_half"]))
    for o, f in zip(out, frames):
        assert np.array_equal(o[:y0], f[:y0]) and np.array_equal(o[y1:], f[y1:])
    sums = np.array([int(o[y0:y1].astype(np.int64).sum()) for o in out])
    assert np.abs(sums - z["strip_sums"]).max() / ((y1 - y0) * W * 3) < 0.05
