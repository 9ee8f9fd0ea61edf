"""Scene-cut detection on the B200: drop-in for `SubtitleDetect.get_scene_div_frame_no` (backend/tools/subtitle_detect.py:158-170),
which runs the vendored PySceneDetect `ContentDetector` over a second full CPU decode of the video in ProPainter mode (main.py:165).

The per-frame work — down-scale by W // 256 (cv2.resize INTER_LINEAR), BGR -> 8-bit HSV (cv2.cvtColor), mean |delta| of the three channels against
the previous frame — is one HBM-bound kernel per frame (csrc/elementwise.cuh `scene_hsv_diff_kernel`) that returns three integer sums; the float
score (content_detector.py:25-35, :183-186), the threshold 27.0 / min_scene_len 15 logic (:200-222) and the scene list -> frame numbers step
(subtitle_detect.py:163-169) run on the host on those integers and are therefore bit-exact.  No CPU fallback."""
import ctypes as C
from typing import Iterable, List, Sequence

import numpy as np

from . import _capi

THRESHOLD, MIN_SCENE_LEN = 27.0, 15      # ContentDetector defaults (content_detector.py:117-118), used as-is by subtitle_detect.py:163


def score_from_sums(sums: Sequence[int], num_pixels: int) -> float:
    """content_detector.py:25-35, :183-186: the three components weigh 1.0, delta_edges 0.0; float64 scalars of numpy added left to right."""
    comps = [np.float64(int(s)) / float(num_pixels) for s in sums] + [0.0]
    weights = (1.0, 1.0, 1.0, 0.0)
    return float(sum(c * w for c, w in zip(comps, weights)) / sum(abs(w) for w in weights))


def cuts_from_scores(scores: Sequence[float], first_frame: int = 0) -> List[int]:
    """process_frame (content_detector.py:188-222)."""
    cuts, last = [], None
    for i, sc in enumerate(scores):
        n = first_frame + i
        if last is None:
            last = n
        if sc >= THRESHOLD and (n - last) >= MIN_SCENE_LEN:
            last = n
            cuts.append(n)
    return cuts


class SceneScorer:
    """Feeds decoded BGR frames to the device in batches and yields the reference's per-frame scores."""

    def __init__(self, device="cuda:0", runtime=None, batch: int = 16):
        from .dbnet import _DeviceRuntime

        self._rt = runtime if runtime is not None else _DeviceRuntime(device)
        self._own = runtime is None
        self.batch = batch
        self._size = None

    def close(self):
        if self._own and self._rt is not None:
            self._rt.close()
        self._rt = None

    def scores(self, frames: Iterable[np.ndarray]) -> List[float]:
        L, h = self._rt.L, self._rt.h
        out: List[float] = []
        buf: List[np.ndarray] = []
        npx = None

        def flush():
            if not buf:
                return
            ptrs = (C.c_void_p * len(buf))(*[f.ctypes.data for f in buf])
            sums = np.zeros((len(buf), 3), np.int64)
            _capi.check(L.vsr_rt_scene_frames(h, C.cast(ptrs, C.POINTER(C.c_void_p)), len(buf), _capi.ptr(sums, C.c_int64)))
            for s in sums:
                out.append(0.0 if not out else score_from_sums(s, npx))
            buf.clear()

        first = True
        for f in frames:
            f = np.ascontiguousarray(f, np.uint8)
            if first:               # a new sequence: fix the size, forget the previous frame
                first = False
                H, W = f.shape[:2]
                _capi.check(L.vsr_rt_scene_begin(h, H, W))
                self._size = (H, W)
                fac = 1 if W < 256 else W // 256
                npx = (H * W) if fac <= 1 else int(round(H / fac)) * int(round(W / fac))
            if f.shape != self._size + (3,):
                raise ValueError("all frames of a video share one [H,W,3] shape")
            buf.append(f)
            if len(buf) >= self.batch:
                flush()
        flush()
        return out


def scene_div_frame_no(frames: Iterable[np.ndarray], device="cuda:0", runtime=None) -> List[int]:
    """`SubtitleDetect.get_scene_div_frame_no` on decoded frames: every scene start c > 0 contributes c + 1 (subtitle_detect.py:163-169)."""
    sc = SceneScorer(device, runtime)
    try:
        return [c + 1 for c in cuts_from_scores(sc.scores(frames)) if c != 0]
    finally:
        sc.close()


def get_scene_div_frame_no(v_path: str, device="cuda:0") -> List[int]:
    """Same signature as the reference's static method: decodes `v_path` with cv2.VideoCapture (video I/O stays on the host)."""
    import cv2

    cap = cv2.VideoCapture(v_path)

    def frames():
        while True:
            ok, f = cap.read()
            if not ok:
                break
            yield f

    try:
        return scene_div_frame_no(frames(), device)
    finally:
        cap.release()
