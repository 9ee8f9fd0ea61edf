"""CPU: the overlapped detection / inpaint loop (vsr_b200/pipeline_async.py) against the synchronous mirror of
`SubtitleRemover.video_inpaint` (vsr_b200/pipeline.py) — intervals, masks and output frames must be identical for any detection pattern."""
import numpy as np
import pytest

from vsr_b200 import subtitle_plan as P
from vsr_b200.pipeline import interval_boxes, plan_intervals, video_inpaint_frames
from vsr_b200.pipeline_async import StreamingPlanner, video_inpaint_frames_overlapped


def _pattern(rng, n, step):
    """random subtitle runs: stretches with a (slightly jittering) box, gaps of random length, isolated single hits"""
    det, f = {}, 1
    while f <= n:
        gap = int(rng.integers(0, 40))
        f += gap
        run = int(rng.choice([1, 1, 2, 5, 12, 30, 70]))
        base = (int(rng.integers(100, 140)), int(rng.integers(500, 560)), int(rng.integers(300, 320)), int(rng.integers(340, 360)))
        for g in range(f, min(f + run, n + 1)):
            if P.is_sampled(g, step) and rng.random() < 0.9:
                j = int(rng.integers(-3, 4))
                boxes = [(base[0] + j, base[1] + j, base[2], base[3])]
                if rng.random() < 0.15:
                    boxes.append((base[0], base[1], base[2] - 80, base[3] - 80))
                det[g] = boxes
        f += run
    return det


@pytest.mark.parametrize("step", [2, 3, 4])
def test_streaming_plan_equals_batch_plan(step):
    rng = np.random.default_rng(100 + step)
    for trial in range(300):
        n = int(rng.integers(5, 400))
        det = _pattern(rng, n, step)
        sub = P.drop_empty(P.unify_regions(P.gap_fill(det, step)))
        plan = plan_intervals(sub, n)
        want = [(s, plan[s], interval_boxes(sub, s, plan[s])) for s in sorted(plan)]
        planner, got = StreamingPlanner(n, step), []
        for no in range(1, n + 1):
            if P.is_sampled(no, step):
                got += planner.feed(no, det.get(no, []))
        early = len(got)
        rest, sub2 = planner.finish()
        got += rest
        assert got == want and sub2 == sub, (trial, n, step)
        if len(want) >= 3:
            assert early >= 1, "nothing became final before the end of a long clip"


class _Detector:
    SAMPLE_STEP = 3

    def __init__(self, det):
        self.det = det

    def detect_subtitle(self, frame):
        return self.det.get(int(frame[0, 0, 0]) * 256 + int(frame[0, 0, 1]), [])

    def scan_frames(self, frames, sections=None, on_frame=None):
        sampled = {}
        for no, f in enumerate(frames, 1):
            if P.is_sampled(no, self.SAMPLE_STEP):
                b = self.detect_subtitle(f)
                if b:
                    sampled[no] = b
        return P.drop_empty(P.unify_regions(P.gap_fill(sampled, self.SAMPLE_STEP)))


def _model(batch, mask):
    return [np.where(mask[:, :, None] > 0, (f.astype(np.int32) * 3 + len(batch)) % 251, f).astype(np.uint8) for f in batch]


def test_overlapped_loop_equals_synchronous_loop():
    rng = np.random.default_rng(7)
    for trial in range(6):
        n = int(rng.integers(40, 260))
        det = _pattern(rng, n, 3)
        frames = []
        for i in range(1, n + 1):
            f = rng.integers(0, 255, (400, 640, 3), dtype=np.uint8)
            f[0, 0, 0], f[0, 0, 1] = i // 256, i % 256
            frames.append(f)
        d = _Detector(det)
        want, wsub, wse = video_inpaint_frames(frames, d, _model)
        got, sub, se = video_inpaint_frames_overlapped(frames, d, _model, queue_depth=4)
        assert sub == wsub and se == wse and len(got) == len(want) == n
        assert all(np.array_equal(a, b) for a, b in zip(got, want)), trial


def test_detector_failure_reaches_the_caller():
    class Boom(_Detector):
        def detect_subtitle(self, frame):
            raise RuntimeError("detector died")

    frames = [np.zeros((8, 8, 3), np.uint8) for _ in range(10)]
    with pytest.raises(RuntimeError, match="detector died"):
        video_inpaint_frames_overlapped(frames, Boom({}), _model)
