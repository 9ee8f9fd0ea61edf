"""In-memory mirror of `SubtitleRemover.video_inpaint` (backend/main.py:260-333), the loop of BASELINE config 4 and of the LAMA /
ProPainter modes: detect -> plan intervals -> one mask per interval -> `model(batch, mask)` per `batch_generator` batch.

The reference reads frames from a cv2.VideoCapture and writes to a VideoWriter; here both ends are Python sequences so the
whole chain can run (and be measured) without video I/O, which SURVEY §2 keeps out of scope.  Planning is bit-exact
(vsr_b200.subtitle_plan); detection and inpainting run on the B200."""
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np

from . import subtitle_plan as P
from .config import config
from .inpaint_tools import batch_generator, create_mask


def plan_intervals(sub_list: Dict[int, List[P.Box]], frame_count: int) -> Dict[int, int]:
    """main.py:265-277: {first frame: last frame} (1-based, inclusive) of the intervals that go through the inpaint model."""
    if not sub_list:
        return {}
    ranges = P.find_continuous_ranges_with_same_mask(sub_list)
    ranges = P.expand_frame_ranges(ranges, config.subtitleTimelineBackwardFrameCount.value, config.subtitleTimelineForwardFrameCount.value)
    ranges = P.filter_and_merge_intervals(ranges, config.sttnReferenceLength.value)
    return {s: min(e, frame_count) for s, e in ranges}


def interval_boxes(sub_list: Dict[int, List[P.Box]], start: int, end: int) -> List[P.Box]:
    """main.py:312-322: union of the boxes of frames start..end-1, minus boxes that are much taller than wide."""
    out: List[P.Box] = []
    for no in range(start, end):
        for area in sub_list.get(no, ()):
            xmin, xmax, ymin, ymax = area
            if (ymax - ymin) - (xmax - xmin) > config.subtitleYXAxisDifferencePixel.value:
                continue
            if area not in out:
                out.append(area)
    return out


def video_inpaint_frames(frames: Sequence[np.ndarray], detector, model: Callable, on_interval=None) -> Tuple[List[np.ndarray], Dict[int, List[P.Box]], Dict[int, int]]:
    """frames: BGR uint8 [H,W,3]; detector: SubtitleDetect-like (`scan_frames`); model: STTNDetInpaint / LamaInpaint-like
    `model(batch, mask) -> frames`.  Returns (output frames, detected frame dictionary, interval map)."""
    n = len(frames)
    sub_list = detector.scan_frames(frames)
    start_end = plan_intervals(sub_list, n)
    size = frames[0].shape[:2]
    out: List[np.ndarray] = []
    i = 1
    while i <= n:
        if i not in start_end:
            out.append(frames[i - 1])
            i += 1
            continue
        s, e = i, start_end[i]
        need = list(frames[s - 1:e])
        mask = create_mask(size, interval_boxes(sub_list, s, e))
        if on_interval is not None:
            on_interval(s, e, mask)
        for batch in batch_generator(need, config.getSttnMaxLoadNum()):
            if len(batch) >= 1:
                out.extend(model(batch, mask))
        i = e + 1
    return out, sub_list, start_end


def propainter_mode_frames(frames: Sequence[np.ndarray], sub_list: Dict[int, List[P.Box]], propainter: Callable, lama, scene_points: Sequence[int] = ()) -> List[np.ndarray]:
    """In-memory mirror of `SubtitleRemover.propainter_mode` (backend/main.py:150-246), the loop of BASELINE configs 3 and 5, from the detected
    frame dictionary on: frames without a detection pass through; an interval of frames with the same mask (`find_continuous_ranges_with_same_mask`,
    split at `scene_points`; the reference finds those with the third-party scenedetect, main.py:165) is cut into `batch_generator` batches of
    at most `propainterMaxLoadNum` frames that go through `propainter(batch, mask)` with the mask of the interval's FIRST frame; a single
    frame — an interval or a batch of one — goes to `lama.inpaint(frame, mask)` instead.  Kept from the reference: in the batch-of-one
    branch it is the LAST frame read of the interval that is inpainted (`frame`, :229-231; batches of one are last batches, so this is
    the frame of the batch), and a detected frame that starts no interval is neither processed nor written (:190 has no else).
    Returns the frames in the order the reference writes them."""
    size = frames[0].shape[:2]
    ranges = P.split_range_by_scene(P.find_continuous_ranges_with_same_mask(sub_list), list(scene_points))
    out: List[np.ndarray] = []
    index, n = 0, len(frames)
    while index < n:
        index += 1
        frame = frames[index - 1]
        if index not in sub_list:
            out.append(frame)
            continue
        if not any(s == index for s, _ in ranges):
            continue
        start = index
        end = next((e for s, e in ranges if s <= index <= e), -1)
        if end == -1:
            continue
        temp = [frame]
        while index < end and index < n:
            index += 1
            frame = frames[index - 1]
            temp.append(frame)
        if len(temp) == 1:
            out.append(lama.inpaint(frame, create_mask(size, sub_list[index])))
            continue
        mask = create_mask(size, sub_list[start])
        for batch in batch_generator(temp, config.propainterMaxLoadNum.value):
            if len(batch) == 1:
                out.append(lama.inpaint(frame, mask))
            elif len(batch) > 1:
                out.extend(propainter(batch, mask))
    return out
