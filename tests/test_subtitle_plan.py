"""CPU: the product's planning functions (vsr_b200.subtitle_plan) against golden vectors
produced by the UNMODIFIED reference (tools/make_golden.py), bit-exact."""
import json
import os

import numpy as np

from conftest import GOLDEN
from vsr_b200 import subtitle_plan as P


def _ikeys(d):
    return {int(k): [tuple(b) for b in v] for k, v in d.items()}


def _pairs(x):
    return [tuple(p) for p in x]


def test_planning_matches_reference_vectors():
    z = np.load(os.path.join(GOLDEN, "subtitle_plan.npz"))
    cases = json.loads(str(z["plan"]))
    assert len(cases) > 30
    for c in cases:
        sampled = _ikeys(c["sampled"])
        filled = P.gap_fill(sampled, c["step"])
        assert filled == _ikeys(c["filled"])
        uni = P.drop_empty(P.unify_regions(dict(filled)))
        assert uni == _ikeys(c["unified"])
        assert P.find_continuous_ranges_with_same_mask(uni) == _pairs(c["ranges"])
        assert P.find_continuous_ranges(uni) == _pairs(c["ranges0"])
        exp = P.expand_frame_ranges(_pairs(c["ranges"]), 3, 3)
        assert exp == _pairs(c["expanded"])
        merged = P.filter_and_merge_intervals(exp, 10)
        assert merged == _pairs(c["merged"])
        assert P.split_range_by_scene(merged, list(c["points"])) == _pairs(c["split"])
    quads = json.loads(str(z["quads"]))
    assert P.get_coordinates(quads) == _pairs(json.loads(str(z["coords"])))


def test_sampling_and_filtering_rules():
    assert [P.sample_step_for_fps(f) for f in (23.976, 25, 29.97, 30, 59.94, 60, 120)] == [2, 2, 2, 3, 3, 4, 4]
    assert [f for f in range(1, 12) if P.is_sampled(f, 3)] == [1, 4, 7, 10]
    boxes = [(300, 900, 960, 1040), (10, 50, 10, 40), (300, 1700, 960, 1040)]
    area = [(950, 1069, 288, 1632)]  # (ymin, ymax, xmin, xmax)
    assert P.filter_boxes(boxes, area) == [boxes[0]]
    assert P.filter_boxes(boxes, None) == boxes and P.filter_boxes(boxes, []) == boxes
    assert P.filter_boxes(boxes, area + [(0, 100, 0, 100)]) == boxes[:2]
    assert P.filter_and_merge_intervals([], 10) == [] and P.expand_frame_ranges([], 3, 3) == []


def test_video_inpaint_frames_loop_on_cpu():
    """The in-memory mirror of SubtitleRemover.video_inpaint (main.py:260-333) with stand-in detector / model: interval map,
    mask per interval, batch_generator batching and pass-through of the frames outside every interval."""
    import numpy as np
    from vsr_b200 import pipeline as PL
    from vsr_b200 import subtitle_plan as P
    from vsr_b200.config import config

    class Det:
        SAMPLE_STEP = 3

        def scan_frames(self, frames):
            sampled = {no: [(100, 400, 300, 330)] for no in range(1, len(frames) + 1, 3) if 20 <= no <= 130}
            return P.drop_empty(P.unify_regions(P.gap_fill(sampled, 3)))

    calls = []

    def model(batch, mask):
        calls.append((len(batch), int(mask[315, 250]), int(mask[10, 10])))
        return [b + 1 for b in batch]

    frames = [np.full((360, 480, 3), i % 200, np.uint8) for i in range(160)]
    out, sub, se = PL.video_inpaint_frames(frames, Det(), model)
    (s, e), = se.items()
    assert (s, e) == (min(sub) - 3, max(sub) + 3) and len(out) == 160          # expanded by 3 frames on both sides
    assert [c[0] for c in calls] == [len(b) for b in PL.batch_generator(list(range(e - s + 1)), config.getSttnMaxLoadNum())]
    assert all(c[1] == 255 and c[2] == 0 for c in calls)
    for i, (o, f) in enumerate(zip(out, frames), 1):
        assert np.array_equal(o, f + 1 if s <= i <= e else f)
    assert PL.plan_intervals({}, 10) == {} and PL.interval_boxes({5: [(0, 10, 0, 100)]}, 5, 6) == []   # tall box dropped


def test_propainter_mode_chain_routes_frames_like_the_reference_loop():
    """main.py:176-246 with stand-in models: pass-through frames, one interval cut into batch_generator batches of at most
    propainterMaxLoadNum frames with the FIRST frame's mask, single frames (an interval or a batch of one) to LAMA."""
    import numpy as np
    from vsr_b200.config import config
    from vsr_b200.inpaint_tools import batch_generator, create_mask
    from vsr_b200.pipeline import propainter_mode_frames

    H, W, n = 64, 96, 40
    frames = [np.full((H, W, 3), i, np.uint8) for i in range(1, n + 1)]          # frame number in every pixel
    box_a, box_b = (10, 60, 40, 50), (30, 80, 20, 30)
    sub = {i: [box_a] for i in range(5, 28)}                                      # frames 5..27: one interval of 23 frames
    sub[33] = [box_b]                                                             # frame 33: an interval of one frame
    calls = []

    class Lama:
        def inpaint(self, frame, mask):
            calls.append(("lama", int(frame[0, 0, 0]), mask.copy()))
            return frame + 100

    def propainter(batch, mask):
        calls.append(("pp", [int(f[0, 0, 0]) for f in batch], mask.copy()))
        return [f + 200 for f in batch]

    saved = config.propainterMaxLoadNum.value
    config.propainterMaxLoadNum.value = 11
    try:
        out = propainter_mode_frames(frames, sub, propainter, Lama())
        sizes = [len(b) for b in batch_generator(list(range(23)), 11)]
    finally:
        config.propainterMaxLoadNum.value = saved
    assert len(out) == n and [int(f[0, 0, 0]) for f in out[:4]] == [1, 2, 3, 4]
    assert all(int(out[i - 1][0, 0, 0]) == (i + 200) % 256 for i in range(5, 28)) and int(out[32][0, 0, 0]) == 133 and int(out[39][0, 0, 0]) == 40
    pp = [c for c in calls if c[0] == "pp"]
    assert [len(c[1]) for c in pp] == sizes and sum(sizes) == 23 and [x for c in pp for x in c[1]] == list(range(5, 28))
    assert all(np.array_equal(c[2], create_mask((H, W), [box_a])) for c in pp)
    lama = [c for c in calls if c[0] == "lama"]
    assert len(lama) == 1 and lama[0][1] == 33 and np.array_equal(lama[0][2], create_mask((H, W), [box_b]))

    # an interval whose last batch has one frame: that frame goes to LAMA with the interval's mask
    calls.clear()
    sub2 = {i: [box_a] for i in range(2, 5)}                                      # 3 frames, batches of at most 2 -> 2 + 1
    config.propainterMaxLoadNum.value = 2
    try:
        out = propainter_mode_frames(frames[:6], sub2, propainter, Lama())
    finally:
        config.propainterMaxLoadNum.value = saved
    assert [c[0] for c in calls] == ["pp", "lama"] and calls[0][1] == [2, 3] and calls[1][1] == 4
    assert [int(f[0, 0, 0]) for f in out] == [1, 202, 203, 104, 5, 6]
    # a scene change inside the interval splits it (subtitle_detect.py:135-156)
    calls.clear()
    config.propainterMaxLoadNum.value = 70
    try:
        propainter_mode_frames(frames, {i: [box_a] for i in range(5, 15)}, propainter, Lama(), scene_points=[9])
    finally:
        config.propainterMaxLoadNum.value = saved
    assert [c[1] for c in calls] == [[5, 6, 7, 8], [9, 10, 11, 12, 13, 14]]
