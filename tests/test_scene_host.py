"""CPU: the host half of the device scene-cut path (vsr_b200/scene_detect.py: score from the kernel's integer sums, threshold / min-scene-length
logic) against the oracle that is pinned to the reference's ContentDetector — so that only the kernel's sums remain for the GPU test."""
import numpy as np

from oracle import scene_oracle as S
from vsr_b200 import scene_detect as D


def test_scores_and_cuts_from_sums_equal_the_oracle():
    rng = np.random.default_rng(4)
    for _ in range(200):
        npx = int(rng.integers(1000, 60000))
        sums = [int(v) for v in rng.integers(0, 255 * npx, 3)]
        assert D.score_from_sums(sums, npx) == S.score_from_sums(sums, npx)
    for _ in range(200):
        scores = list(rng.choice([0.0, 3.0, 26.999, 27.0, 80.0], int(rng.integers(1, 120))))
        first = int(rng.integers(0, 5))
        assert D.cuts_from_scores(scores, first) == S.cuts_from_scores(scores, first)
    assert (D.THRESHOLD, D.MIN_SCENE_LEN) == (S.THRESHOLD, S.MIN_SCENE_LEN) == (27.0, 15)


def test_downscale_size_matches_the_entry_point_rule():
    """vsr_rt_scene_begin computes W // 256, round-half-even sizes and the 2x2-mean mode in C; this is the same rule in the oracle"""
    for (H, W), want in {(1080, 1920): (7, 154, 274), (720, 1280): (5, 144, 256), (360, 640): (2, 180, 320), (480, 852): (3, 160, 284),
                         (2160, 3840): (15, 144, 256), (200, 250): (1, 200, 250), (271, 600): (2, 136, 300), (1, 256): (1, 1, 256)}.items():
        f, h, w = S.downscale_size(H, W)
        assert (f, h, w) == (want[0], want[1], want[2]) or (H, W) == (1, 256)
