"""Overlapped form of `SubtitleRemover.video_inpaint` (backend/main.py:260-333) — SURVEY.md §8 (f-2): the reference first detects the whole
video (`find_subtitle_frame_no`), then plans, then inpaints interval by interval.  Detection (DBNet on the graph runtime's stream) and
inpainting (STTN-det / LAMA on the engine's stream) are independent GPU work, so here a producer thread detects the sampled frames in order
while the caller's thread inpaints every interval as soon as the plan for it can no longer change.  The output is the synchronous loop's,
frame for frame (tests/test_pipeline_async.py: random detection patterns against `pipeline.video_inpaint_frames`).

Why an interval becomes final (all functions are forward scans, vsr_b200/subtitle_plan.py):
  * gap filling (subtitle_detect.py:112-124) joins two sampled hits at most 2*step apart: once the detection frontier F is 2*step past the
    last hit h, no later hit reaches back; the filled dictionary is final up to K = F (else up to K = h);
  * `unify_regions` (:181-215) walks the keys forwards: values at keys <= K only depend on keys <= K;
  * `find_continuous_ranges_with_same_mask`, `expand_frame_ranges` (+-3 frames, clipped at the neighbours) and `filter_and_merge_intervals`
    (growth by 4, merging while touching and short) look at most one interval ahead: anything that starts after K can begin no earlier
    than K + 1 - 3 - 4 once widened, so an interval of the plan of the prefix that ends more than MARGIN frames before K is the interval of the
    whole video's plan."""
import queue
import threading
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np

from . import subtitle_plan as P
from .config import config
from .inpaint_tools import batch_generator, create_mask
from .pipeline import interval_boxes, plan_intervals

MARGIN = 12   # frames between an interval's end and the final frontier K: 3 (forward widening) + 3 (backward widening of a later range) + 4 (growth) + 2


class StreamingPlanner:
    """Detections of the sampled frames arrive in increasing frame number; `feed` returns the intervals (start, end, boxes) that have just
    become final, `finish` the remaining ones.  The plan of every prefix is computed with the very functions of the synchronous path."""

    def __init__(self, n_frames: int, step: int):
        self.n, self.step = n_frames, step
        self.sampled: Dict[int, List[P.Box]] = {}
        self.last_hit = 0
        self.emitted_upto = 0           # every frame <= this belongs to an emitted interval or to none

    def _plan(self, upto: int):
        sub = P.drop_empty(P.unify_regions(P.gap_fill(self.sampled, self.step)))
        sub = {k: v for k, v in sub.items() if k <= upto}
        return sub, plan_intervals(sub, self.n)

    def _emit(self, final_upto: int, everything: bool):
        sub, plan = self._plan(self.n if everything else final_upto)
        out = []
        for s in sorted(plan):
            e = plan[s]
            if s <= self.emitted_upto:
                continue
            if not everything and e + MARGIN >= final_upto:
                break                   # this one and everything after it can still change
            out.append((s, e, interval_boxes(sub, s, e)))
            self.emitted_upto = e
        return out, sub

    def feed(self, frame_no: int, boxes: Sequence[P.Box]):
        """`frame_no` = 1-based number of the sampled frame just detected (strictly increasing), `boxes` its detections (may be empty)."""
        if boxes:
            self.sampled[frame_no] = list(boxes)
            self.last_hit = frame_no
        K = frame_no if frame_no - self.last_hit >= 2 * self.step or not self.sampled else self.last_hit
        return self._emit(K, False)[0]

    def finish(self):
        out, sub = self._emit(self.n, True)
        return out, sub


def video_inpaint_frames_overlapped(frames: Sequence[np.ndarray], detector, model: Callable, queue_depth: int = 64):
    """Same result as `pipeline.video_inpaint_frames(frames, detector, model)`: (output frames, detected frame dictionary, interval map).
    `detector` needs `detect_subtitle(frame)` and `SAMPLE_STEP`; `model(batch, mask)` is STTNDetInpaint / LamaInpaint-like."""
    n = len(frames)
    step = detector.SAMPLE_STEP
    q: "queue.Queue" = queue.Queue(maxsize=queue_depth)
    failure: List[BaseException] = []

    def detect():
        try:
            for no in range(1, n + 1):
                if P.is_sampled(no, step):
                    q.put((no, detector.detect_subtitle(frames[no - 1])))
        except BaseException as e:      # noqa: BLE001 — re-raised in the caller's thread
            failure.append(e)
        finally:
            q.put(None)

    producer = threading.Thread(target=detect, name="vsr-detect", daemon=True)
    producer.start()
    size = frames[0].shape[:2]
    out: List[np.ndarray] = list(frames)
    start_end: Dict[int, int] = {}
    planner = StreamingPlanner(n, step)

    def inpaint(s, e, boxes):
        start_end[s] = e
        mask = create_mask(size, boxes)
        pos = s - 1
        for batch in batch_generator(list(frames[s - 1:e]), config.getSttnMaxLoadNum()):
            if len(batch) >= 1:
                out[pos:pos + len(batch)] = model(batch, mask)
                pos += len(batch)

    while True:
        item = q.get()
        if item is None:
            break
        for s, e, boxes in planner.feed(*item):
            inpaint(s, e, boxes)
    producer.join()
    if failure:
        raise failure[0]
    rest, sub_list = planner.finish()
    for s, e, boxes in rest:
        inpaint(s, e, boxes)
    return out, sub_list, start_end
