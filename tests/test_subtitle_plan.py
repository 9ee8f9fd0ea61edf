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
