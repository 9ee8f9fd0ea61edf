"""Integer planning path around the text detector: which frames get
detected, how sampled hits are gap-filled and unified, and how hit frames become inpaint intervals.
Pure host logic, bit-exact against the reference functions it mirrors (tests/test_subtitle_plan.py runs the
golden vectors produced by the unmodified reference, and the reference itself when it is present).

Keys of the frame dictionaries are 1-based frame numbers, boxes are (xmin, xmax, ymin, ymax) ints, exactly as
in backend/tools/subtitle_detect.py.
"""
from typing import Dict, Iterable, List, Sequence, Tuple

from .config import config

Box = Tuple[int, int, int, int]


def sample_step_for_fps(fps: float) -> int:
    """SubtitleDetect._init_sample_step (subtitle_detect.py:29-39): >= 8 detections per second."""
    if fps >= 60:
        return 4
    return 3 if fps >= 30 else 2


def is_sampled(frame_no_1based: int, step: int) -> bool:
    """subtitle_detect.py:105: frames 1, 1+step, 1+2*step, ... are detected."""
    return step <= 1 or (frame_no_1based - 1) % step == 0


def get_coordinates(dt_polys: Iterable) -> List[Box]:
    """backend/tools/ocr.py:1-20: quad [tl, tr, br, bl] -> the axis box inscribed between its corners."""
    out = []
    for quad in dt_polys:
        xs = [int(p[0]) for p in quad]
        ys = [int(p[1]) for p in quad]
        out.append((max(xs[0], xs[3]), min(xs[1], xs[2]), max(ys[0], ys[1]), min(ys[2], ys[3])))
    return out


def filter_boxes(boxes: Iterable[Box], sub_areas) -> List[Box]:
    """subtitle_detect.py:60-82: with selection areas (ymin, ymax, xmin, xmax), keep the boxes lying entirely
    inside one of them (each box at most once); without areas keep everything."""
    boxes = list(boxes)
    if not sub_areas:
        return boxes

    def inside(b, a):
        return a[2] <= b[0] and b[1] <= a[3] and a[0] <= b[2] and b[3] <= a[1]

    return [b for b in boxes if any(inside(b, a) for a in sub_areas)]


def gap_fill(sampled: Dict[int, List[Box]], step: int) -> Dict[int, List[Box]]:
    """subtitle_detect.py:112-124: frames strictly between two sampled hits at most 2*step apart inherit the
    boxes of the earlier hit."""
    hits = sorted(sampled)
    out: Dict[int, List[Box]] = {}
    for a, b in zip(hits, hits[1:]):
        out[a] = sampled[a]
        if b - a <= 2 * step:
            for f in range(a + 1, b):
                out[f] = sampled[a]
    if hits:
        out[hits[-1]] = sampled[hits[-1]]
    return out


def _similar(r1: Box, r2: Box) -> bool:
    """SubtitleDetect.are_similar (subtitle_detect.py:173-179)."""
    tx = config.subtitleAreaPixelToleranceXPixel.value
    ty = config.subtitleAreaPixelToleranceYPixel.value
    return abs(r1[0] - r2[0]) <= tx and abs(r1[1] - r2[1]) <= tx and abs(r1[2] - r2[2]) <= ty and abs(r1[3] - r2[3]) <= ty


def unify_regions(raw: Dict[int, List[Box]]) -> Dict[int, List[Box]]:
    """subtitle_detect.py:181-215: walking the keys in order, the i-th box of a frame snaps to the i-th
    (already unified) box of the previous key when the two are within the pixel tolerances."""
    if not raw:
        return raw
    keys = sorted(raw)
    out = {keys[0]: raw[keys[0]]}
    prev = out[keys[0]]
    for k in keys[1:]:
        cur = []
        for i, box in enumerate(raw[k]):
            ref = prev[i] if i < len(prev) else None
            cur.append(ref if ref and _similar(box, ref) else box)
        out[k] = cur
        prev = cur
    return out


def drop_empty(d: Dict[int, List[Box]]) -> Dict[int, List[Box]]:
    """subtitle_detect.py:128-132."""
    return {k: v for k, v in d.items() if len(v) > 0}


def find_continuous_ranges(frames: Dict[int, List[Box]]) -> List[Tuple[int, int]]:
    """subtitle_detect.py:217-236: maximal runs of consecutive frame numbers."""
    nums = sorted(frames)
    out, start = [], nums[0]
    for a, b in zip(nums, nums[1:]):
        if b - a != 1:
            out.append((start, a))
            start = b
    out.append((start, nums[-1]))
    return out


def find_continuous_ranges_with_same_mask(frames: Dict[int, List[Box]]) -> List[Tuple[int, int]]:
    """subtitle_detect.py:238-259: runs of consecutive frame numbers whose box lists are identical."""
    nums = sorted(frames)
    out, start = [], nums[0]
    for a, b in zip(nums, nums[1:]):
        if b - a != 1 or frames[b] != frames[a]:
            out.append((start, a))
            start = b
    out.append((start, nums[-1]))
    return out


def filter_and_merge_intervals(intervals: Sequence[Tuple[int, int]], target_length: int) -> List[Tuple[int, int]]:
    """subtitle_detect.py:261-293: single-frame intervals grow to ~target_length without touching their
    neighbours, then overlapping/adjacent intervals merge while either side is shorter than target_length."""
    if not intervals:
        return []
    iv = sorted(intervals, key=lambda x: x[0])
    half = (target_length - 1) // 2
    grown: List[Tuple[int, int]] = []
    for i, (s, e) in enumerate(iv):
        if s != e:
            grown.append((s, e))
            continue
        lo = s - half
        if grown:
            lo = max(lo, grown[-1][1] + 1)
        hi = s + half
        if i + 1 < len(iv):
            hi = min(hi, iv[i + 1][0] - 1)
        grown.append((lo, hi) if hi >= lo else (s, s))
    merged = [grown[0]]
    for s, e in grown[1:]:
        ls, le = merged[-1]
        touching = s <= le + 1
        if touching and (e - s + 1 < target_length or le - ls + 1 < target_length):
            merged[-1] = (ls, max(le, e))
        else:
            merged.append((s, e))
    return merged


def expand_frame_ranges(ranges: Sequence[Tuple[int, int]], backward: int, forward: int) -> List[Tuple[int, int]]:
    """backend/tools/inpaint_tools.py:244-301: widen every range by (backward, forward) frames without
    overlapping its neighbours; frame numbers stay >= 1."""
    if not ranges:
        return []
    rs = sorted(ranges)
    out: List[Tuple[int, int]] = []
    for i, (s, e) in enumerate(rs):
        ns = max(1, s - backward)
        ne = e + forward
        if i + 1 < len(rs):
            nxt = rs[i + 1][0]
            if ne >= nxt:
                ne = e if nxt - e == 1 else min(ne, nxt - 1)
        if out and ns <= out[-1][1]:
            ns = out[-1][1] + 1
        out.append((ns, ne) if ns <= ne else (s, e))
    return out


def split_range_by_scene(intervals: Sequence[Tuple[int, int]], points: List[int]) -> List[Tuple[int, int]]:
    """subtitle_detect.py:135-156: cut every interval at the scene-change frame numbers that fall inside it."""
    pts = sorted(points)
    out = []
    for s, e in intervals:
        for p in pts:
            if s <= p <= e:
                if s < p:
                    out.append((s, p - 1))
                s = p
        out.append((s, e))
    return out
