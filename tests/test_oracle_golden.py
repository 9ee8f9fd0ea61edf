"""CPU: the oracle (oracle/sttn_oracle.py) against the committed golden vectors, which were produced
by the UNMODIFIED reference (tools/make_golden.py).  These pin the oracle; they run without a GPU."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import sttn_oracle as O


def _strip(seed, T):
    return [O.cv2_resize_linear_u8(f, 640, 120) for f in O.synthetic_clip(T, 360, 1920, seed=seed)]


def test_mask_index_vectors_match_reference():
    z = np.load(os.path.join(GOLDEN, "mask_index.npz"))
    cases = json.loads(str(z["cases"]))
    for c in cases:
        m = O.create_mask((c["H"], c["W"]), [tuple(b) for b in c["boxes"]])
        assert int(m.astype(np.int64).sum()) == c["mask_sum"]
        assert np.flatnonzero(m.any(1))[[0, -1]].tolist() == c["mask_rows"]
        h = int(c["W"] * 3 / 16)
        m01 = (m > 127).astype(np.uint8)
        assert [list(a) for a in O.get_inpaint_area_by_mask(c["W"], c["H"], h, m01)] == c["areas"]
        assert [list(a) for a in O.get_inpaint_area_by_mask(c["W"], c["H"], h, m01, multiple=8)] == c["areas8"]
    for key, sizes in json.loads(str(z["batches"])).items():
        n, mb = (int(v) for v in key.split("@"))
        assert [b - a for a, b in O.batch_generator(n, mb)] == sizes


def test_window_schedule_counts():
    # T=50 -> 10 windows of 10,14,15,14,15,14,15,14,15,14 frames, 104 decoded
    s = O.window_schedule(50)
    assert [len(a) + len(b) for a, b in s] == [10, 14, 15, 14, 15, 14, 15, 14, 15, 14]
    assert sum(len(a) for a, _ in s) == 104


def test_strip_random_weights_matches_reference_network():
    z = np.load(os.path.join(GOLDEN, "sttn_auto_strip_rand.npz"))
    comps = O.inpaint_strip(O.random_weights(int(z["wseed"])), _strip(int(z["seed"]), int(z["T"])))
    got = np.stack([c.astype(np.float32) for c in comps])
    assert np.array_equal(np.array([c.dtype == np.uint8 for c in comps]), z["once"])
    # same fp32 ops, same library: bit-identical in the build container, a quantisation flip at most elsewhere
    assert np.abs(got - z["comps"]).max() <= 1.0
    assert (got != z["comps"]).mean() < 1e-3


def test_resize_restatement_is_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for (sw, sh, dw, dh) in [(1920, 360, 640, 120), (852, 159, 640, 120), (640, 120, 1920, 360), (640, 120, 852, 159)]:
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        assert np.array_equal(cv2.resize(src, (dw, dh)), O.cv2_resize_linear_u8(src, dw, dh))
    # float path: bit-exact with OpenCV's own C++ implementation (IPP builds differ by <= 0.0075)
    was = cv2.useOptimized()
    try:
        cv2.setUseOptimized(False)
        sf = (rng.integers(0, 512, (120, 640, 3)) / 2).astype(np.float32)
        assert np.array_equal(cv2.resize(sf, (1920, 360)), O.cv2_resize_linear_f32(sf, 1920, 360))
    finally:
        cv2.setUseOptimized(was)


@pytest.mark.slow
def test_strip_real_weights_matches_reference(real_weights_path):
    z = np.load(os.path.join(GOLDEN, "sttn_auto_strip_real.npz"))
    w = O.load_weights(real_weights_path)
    taps = {}
    comps = O.inpaint_strip(w, _strip(int(z["seed"]), int(z["T"])), taps=taps)
    got = np.stack([c.astype(np.float32) for c in comps])
    assert np.abs(got - z["comps"]).max() <= 1.0
    assert (got != z["comps"]).mean() < 1e-3
    np.testing.assert_allclose(taps["encoder"][0, :, ::3, ::8].numpy(), z["encoder_f0"], atol=1e-4)


@pytest.mark.slow
def test_call_real_weights_matches_reference(real_weights_path):
    z = np.load(os.path.join(GOLDEN, "sttn_auto_call_real.npz"))
    H, W, T = int(z["H"]), int(z["W"]), int(z["T"])
    frames = O.synthetic_clip(T, H, W, seed=int(z["seed"]))
    mask = O.default_mask(H, W)
    out = O.sttn_call(O.load_weights(real_weights_path), frames, mask)
    y0, y1 = z["areas"][0][:2]
    got = np.stack([o[y0:y1] for o in out])
    want = z["strip_out"].copy()
    # golden `strip_out` came from an IPP-enabled cv2 (float up-scale differs by <=0.0075 before the
    # truncation); `plain_*` patches it to OpenCV's own C++ path, which the oracle restates bit-exactly.
    full = np.stack(out).copy()
    ref_full = np.stack(frames).copy()
    ref_full[:, y0:y1] = want
    flat = ref_full.reshape(-1)
    flat[z["plain_diff_idx"]] = z["plain_diff_val"]
    d = np.abs(full.astype(np.int32) - flat.reshape(full.shape).astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 1e-4
    d_ipp = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert d_ipp.max() <= 1 and (d_ipp > 0).mean() < 0.06
    for o, f in zip(out, frames):  # rows outside the strip are untouched
        assert np.array_equal(o[:y0], f[:y0]) and np.array_equal(o[y1:], f[y1:])


@pytest.mark.slow
def test_det_oracle_matches_reference():
    """STTN-det (D1-D3): oracle vs the unmodified reference's comps and final strips (plain-C++ cv2 path)."""
    from oracle import sttn_det_oracle as D

    p = os.path.join(os.path.dirname(GOLDEN), "..", "weights", "sttn-det", "sttn.pth")
    if not os.path.exists(p):
        pytest.skip("sttn-det checkpoint not staged under weights/")
    z = np.load(os.path.join(GOLDEN, "sttn_det_real.npz"))
    H, W, T = int(z["H"]), int(z["W"]), int(z["T"])
    w = O.load_weights(p)
    frames = O.synthetic_clip(T, H, W, seed=int(z["seed"]))
    mask = O.default_mask(H, W)
    assert D.split_height(H, W) == int(W * 5 / 18)
    y0, y1 = z["areas"][0][:2]
    scaled = [O.cv2_resize_linear_u8(np.ascontiguousarray(f[y0:y1]), 432, 240) for f in frames]
    msmall = O.cv2_resize_linear_u8(np.ascontiguousarray(mask[y0:y1]), 432, 240)
    assert np.array_equal(msmall, z["mask_small"])
    comps = D.inpaint_strip(w, scaled, [msmall] * T)
    got = np.stack([c.astype(np.float32) for c in comps])
    assert np.abs(got - z["comps"]).max() <= 1.0 and (got != z["comps"]).mean() < 1e-3
    out = D.det_call(w, frames, mask)
    d = np.abs(np.stack([o[y0:y1] for o in out]).astype(np.int32) - z["strip_out_plain"].astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3
