"""TEST INFRASTRUCTURE — CPU restatement of the scene-cut detection the reference runs in ProPainter mode
(`SubtitleDetect.get_scene_div_frame_no`, backend/tools/subtitle_detect.py:158-170 -> the vendored PySceneDetect
`scene_detect(v_path, ContentDetector())`).  SURVEY.md §8(f-3).  Integer path: bit-exact.

What the reference computes per decoded frame (backend/scenedetect/):
  scene_manager.py:852-860, 929-933   auto-downscale: factor = W // 256 (1 below 256), frame -> cv2.resize(frame, (round(W/f), round(H/f)), INTER_LINEAR)
  detectors/content_detector.py:160   cv2.cvtColor(BGR2HSV) on uint8 (H in [0,180), fixed-point tables), split
  :25-35, :176-186                    mean |delta| of hue, saturation, value against the previous frame; score = (dh + ds + dv) / 3
                                      (delta_edges has weight 0 and is only computed with a stats manager)
  :200-222                            cut at frame n when score >= 27.0 and n - last_cut >= 15 (last_cut starts at the first frame number)
  scene_manager.py get_scene_list + subtitle_detect.py:163-169: every scene that does not start at frame 0 contributes start + 1
Parity pinned: tests/test_scene_oracle.py checks the HSV conversion and the resize against cv2 itself, and the whole function against the
unmodified reference classes where the reference tree (or baseline/_ref) is present; golden scores / cuts in tests/golden/scene_cuts.npz."""
from typing import List, Sequence, Tuple

import numpy as np

from .sttn_oracle import cv2_resize_linear_u8

THRESHOLD, MIN_SCENE_LEN, MIN_WIDTH = 27.0, 15, 256
_SHIFT = 12


def downscale_size(H: int, W: int) -> Tuple[int, int, int]:
    """(factor, height, width) of the frames the detector sees (scene_manager.py:132-149, 929-933; Python's round = half to even)."""
    f = 1 if W < MIN_WIDTH else W // MIN_WIDTH
    if f <= 1:
        return 1, H, W
    return f, int(round(H / f)), int(round(W / f))


def downscale(frame: np.ndarray) -> np.ndarray:
    """cv2.resize(frame, (w, h), INTER_LINEAR) as the decode thread calls it.  OpenCV turns an exact 2x2 down-scale into INTER_AREA
    (imgproc/resize.cpp: `interpolation == INTER_LINEAR && is_area_fast && iscale_x == 2 && iscale_y == 2`): the rounded mean of 4 pixels."""
    H, W = frame.shape[:2]
    f, h, w = downscale_size(H, W)
    if f == 1:
        return frame
    if W == 2 * w and H == 2 * h:
        s = frame.astype(np.int32)
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    return cv2_resize_linear_u8(frame, w, h)


def _tables():
    i = np.arange(1, 256, dtype=np.float64)
    sdiv = np.zeros(256, np.int64)
    hdiv = np.zeros(256, np.int64)
    sdiv[1:] = np.rint((255 << _SHIFT) / i).astype(np.int64)            # saturate_cast<int>((255 << hsv_shift) / (1. * i))
    hdiv[1:] = np.rint((180 << _SHIFT) / (6.0 * i)).astype(np.int64)    # saturate_cast<int>((180 << hsv_shift) / (6. * i))
    return sdiv, hdiv


_SDIV, _HDIV = _tables()


def bgr_to_hsv_u8(img: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(img, COLOR_BGR2HSV) for uint8 (imgproc/color_hsv: RGB2HSV_b, hrange 180, 12-bit fixed point)."""
    b, g, r = (img[..., k].astype(np.int64) for k in range(3))
    v = np.maximum(np.maximum(b, g), r)
    vmin = np.minimum(np.minimum(b, g), r)
    diff = v - vmin
    vr = v == r
    vg = v == g
    s = (diff * _SDIV[v] + (1 << (_SHIFT - 1))) >> _SHIFT
    h = np.where(vr, g - b, np.where(vg, b - r + 2 * diff, r - g + 4 * diff))
    h = (h * _HDIV[diff] + (1 << (_SHIFT - 1))) >> _SHIFT
    h = h + np.where(h < 0, 180, 0)
    return np.stack([h, s, v], -1).astype(np.uint8)


def frame_sums(hsv: np.ndarray, prev: np.ndarray) -> Tuple[int, int, int]:
    d = np.abs(hsv.astype(np.int32) - prev.astype(np.int32)).reshape(-1, 3).sum(0)
    return int(d[0]), int(d[1]), int(d[2])


def score_from_sums(sums: Sequence[int], num_pixels: int) -> float:
    """content_detector.py:25-35, :183-186 in the same float64 operation order: sum(component * 1.0) / 3.0 with component = sum / n."""
    # numpy float64 scalars on purpose: CPython >= 3.12 sums exact `float`s with Neumaier compensation, numpy scalars (what the reference's
    # components are) with plain left-to-right additions
    comps = [np.float64(s) / float(num_pixels) for s in sums]
    return float(sum(c * w for c, w in zip(comps + [0.0], (1.0, 1.0, 1.0, 0.0))) / sum(abs(w) for w in (1.0, 1.0, 1.0, 0.0)))


def cuts_from_scores(scores: Sequence[float], first_frame: int = 0) -> List[int]:
    """process_frame (:188-222): frame numbers at which a cut is declared (score of frame n compares n with n-1; frame `first_frame` scores 0)."""
    cuts, last = [], None
    for i, sc in enumerate(scores):
        n = first_frame + i
        if last is None:
            last = n
        if sc >= THRESHOLD and (n - last) >= MIN_SCENE_LEN:
            last = n
            cuts.append(n)
    return cuts


def frame_scores(frames: Sequence[np.ndarray]) -> List[float]:
    out, prev = [], None
    for f in frames:
        hsv = bgr_to_hsv_u8(downscale(np.ascontiguousarray(f)))
        out.append(0.0 if prev is None else score_from_sums(frame_sums(hsv, prev), hsv.shape[0] * hsv.shape[1]))
        prev = hsv
    return out


def scene_div_frame_no(frames: Sequence[np.ndarray]) -> List[int]:
    """get_scene_div_frame_no (subtitle_detect.py:158-170): scenes are [0, c1), [c1, c2), ...; every scene start c > 0 gives c + 1."""
    return [c + 1 for c in cuts_from_scores(frame_scores(frames)) if c != 0]
