"""CPU: the scene-cut oracle (oracle/scene_oracle.py) against cv2 itself, the golden scores / cuts of the unmodified vendored
PySceneDetect (tests/golden/scene_cuts.npz, tools/make_golden_scene.py) and, where the reference modules are present, the reference classes."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import ref_import
from oracle import scene_oracle as S

sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_hsv_and_downscale_equal_cv2():
    import cv2

    rng = np.random.default_rng(1)
    cols = rng.integers(0, 256, (400000, 1, 3), dtype=np.uint8)
    cols[:256, 0] = np.arange(256, dtype=np.uint8)[:, None]                     # greys
    assert np.array_equal(S.bgr_to_hsv_u8(cols), cv2.cvtColor(cols, cv2.COLOR_BGR2HSV))
    for H, W in ((1080, 1920), (720, 1280), (360, 640), (480, 852), (200, 250)):
        f = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        fac, h, w = S.downscale_size(H, W)
        want = f if fac == 1 else cv2.resize(f, (round(W / fac), round(H / fac)), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(S.downscale(f), want), (H, W)


def test_golden_scores_and_cuts():
    from make_golden_scene import clip

    z = np.load(os.path.join(GOLDEN, "scene_cuts.npz"))
    assert bool(z["whole_path_equal"])       # at generation time the oracle equalled get_scene_div_frame_no on the encoded clip
    for name in ("a", "b"):
        H, W = (int(v) for v in z[f"{name}_size"])
        frames = clip(H=H, W=W)
        scores = S.frame_scores(frames)
        assert np.array_equal(np.array(scores), z[f"{name}_scores"])            # float64, bit for bit
        assert S.cuts_from_scores(scores) == z[f"{name}_cuts"].tolist() == [20, 45]
        assert S.scene_div_frame_no(frames) == [21, 46]


@pytest.mark.skipif(not ref_import.available(), reason="reference modules not present (neither /root/reference nor baseline/_ref)")
def test_against_the_reference_detector():
    import cv2
    from make_golden_scene import clip

    ref_import.install()
    from backend.scenedetect.detectors import ContentDetector

    frames = clip(seed=9, n=50, H=271, W=600)           # factor 2 with an odd height: the general bilinear path, not the 2x2 mean
    det, scores, cuts = ContentDetector(), [], []
    for i, fr in enumerate(frames):
        small = cv2.resize(fr, (round(600 / 2), round(271 / 2)), interpolation=cv2.INTER_LINEAR)
        cuts += det.process_frame(i, small)
        scores.append(det._frame_score)
    assert np.array_equal(np.array(S.frame_scores(frames)), np.array(scores)) and S.cuts_from_scores(scores) == cuts
