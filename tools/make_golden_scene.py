#!/usr/bin/env python
"""tests/golden/scene_cuts.npz: the UNMODIFIED vendored PySceneDetect of the reference (backend/scenedetect) on a synthetic clip with hard cuts.
Two ways, both stored: (1) `SubtitleDetect.get_scene_div_frame_no(path)` (subtitle_detect.py:158-170) on the clip written to an .mp4
(the whole reference path: video stream, auto-downscale, detector, scene list); (2) the reference's ContentDetector fed frame by frame with
the frames decoded from that same file and down-scaled exactly as scene_manager.py:929-933 does — per-frame scores.  The test rebuilds the
clip from its seed, so only scores and cut lists are stored.  Build container only."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, scene_oracle as S, sttn_oracle as O  # noqa: E402


def clip(seed=5, n=70, H=360, W=640):
    """scenes of 20 / 25 / 8 / 17 frames: one texture per scene with a little per-frame noise and a brightness drift (scores of a few units
    inside a scene, far above the threshold across a cut; the 8-frame scene is shorter than min_scene_len, so its start is not a cut)"""
    lens, frames = [20, 25, 8, 17], []
    rng = np.random.default_rng(seed)
    for si, ln in enumerate(lens):
        base = O.synthetic_clip(1, H, W, seed=seed + 10 * si)[0].astype(np.int32)
        tint = np.array([(40 * si) % 90, (70 * si) % 120, (25 * si) % 60], np.int32)
        for k in range(ln):
            noise = rng.integers(-3, 4, base.shape)
            frames.append(np.clip(base // 2 + tint + 30 * si + k + noise, 0, 255).astype(np.uint8))
    return frames[:n]


def decode(path):
    import cv2

    cap, out = cv2.VideoCapture(path), []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        out.append(f)
    return out


def write_video(frames, path, fps=25.0):
    import cv2

    w = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (frames[0].shape[1], frames[0].shape[0]))
    for f in frames:
        w.write(f)
    w.release()


def main():
    import cv2

    ref_import.install()
    from backend.scenedetect.detectors import ContentDetector
    from backend.scenedetect.scene_manager import compute_downscale_factor
    from backend.tools.subtitle_detect import SubtitleDetect

    out = {}
    for name, (H, W) in (("a", (360, 640)), ("b", (480, 852))):
        frames = clip(H=H, W=W)
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "clip.mp4")
            write_video(frames, path)
            div = SubtitleDetect.get_scene_div_frame_no(path)
            dec = decode(path)
        det = ContentDetector()
        f = compute_downscale_factor(W)
        scores = []
        for i, fr in enumerate(dec):
            small = cv2.resize(fr, (round(W / f), round(H / f)), interpolation=cv2.INTER_LINEAR) if f > 1 else fr
            det.process_frame(i, small)
            scores.append(det._frame_score)
        mine = S.scene_div_frame_no(dec)
        print(name, (H, W), "reference:", div, "oracle on the decoded frames:", mine, "frames", len(dec))
        assert mine == div and np.array_equal(np.array(S.frame_scores(dec)), np.array(scores))
        # the same functions on the un-encoded frames (what the tests can rebuild from the seed)
        det = ContentDetector()
        raw_scores, raw_cuts = [], []
        for i, fr in enumerate(frames):
            small = cv2.resize(fr, (round(W / f), round(H / f)), interpolation=cv2.INTER_LINEAR) if f > 1 else fr
            raw_cuts += det.process_frame(i, small)
            raw_scores.append(det._frame_score)
        out[f"{name}_size"] = np.array([H, W])
        out[f"{name}_scores"] = np.array(raw_scores, np.float64)
        out[f"{name}_cuts"] = np.array(raw_cuts, np.int64)
        out[f"{name}_div_from_file"] = np.array(div, np.int64)
    p = os.path.join(ROOT, "tests", "golden", "scene_cuts.npz")
    np.savez_compressed(p, whole_path_equal=True, **out)
    print(p, os.path.getsize(p), {k: v.tolist() for k, v in out.items() if "cuts" in k or "div" in k})


if __name__ == "__main__":
    main()
