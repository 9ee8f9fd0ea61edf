"""GPU: scene-cut scores — the device kernel's integer sums, the scores and the cut list against the oracle
(oracle/scene_oracle.py, pinned to cv2 and to the reference's vendored ContentDetector).  Integer path: bit-exact."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import scene_oracle as S

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("H,W", [(360, 640), (480, 852), (200, 250), (271, 600)])
def test_scores_and_cuts_equal_oracle(capi, H, W):
    """2x2-mean down-scale (640x360), bilinear down-scale (852x480, 600x271), no down-scale (250x200)"""
    from make_golden_scene import clip
    from vsr_b200.scene_detect import SceneScorer, cuts_from_scores, scene_div_frame_no

    if capi.lib().vsr_device_count() < 1:
        pytest.fail("GPU tests need a B200 (sm_100) device")
    frames = clip(H=H, W=W)
    sc = SceneScorer("cuda:0", batch=7)
    got = sc.scores(frames)
    assert np.array_equal(np.array(got), np.array(S.frame_scores(frames)))          # float64 scores, bit for bit
    assert cuts_from_scores(got) == S.cuts_from_scores(got) == [20, 45]
    again = sc.scores(frames[30:])                                                   # a second sequence on the same scorer starts afresh
    sc.close()
    assert again[0] == 0.0 and np.array_equal(np.array(again), np.array(S.frame_scores(frames[30:])))
    assert scene_div_frame_no(frames) == S.scene_div_frame_no(frames) == [21, 46]


def test_golden_of_the_reference_and_1080p(capi):
    from make_golden_scene import clip
    from vsr_b200.scene_detect import SceneScorer

    z = np.load(os.path.join(GOLDEN, "scene_cuts.npz"))
    sc = SceneScorer("cuda:0")
    for name in ("a", "b"):
        H, W = (int(v) for v in z[f"{name}_size"])
        assert np.array_equal(np.array(sc.scores(clip(H=H, W=W))), z[f"{name}_scores"])   # the unmodified ContentDetector's scores
    rng = np.random.default_rng(3)
    big = [rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8) for _ in range(3)]        # factor 7 -> 274 x 154
    assert np.array_equal(np.array(sc.scores(big)), np.array(S.frame_scores(big)))
    sc.close()
