"""CPU: the whole ProPainter device pipeline (vsr_b200.propainter_inpaint.PropainterInpaint: RAFT -> flow completion -> image
propagation -> generator per window -> composite) on the fp32 stand-in of the runtime against the frames of the UNMODIFIED
reference (tests/golden/propainter_real.npz `comp`)."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
DIR = os.path.join(ROOT, "weights", "propainter")
pytestmark = pytest.mark.skipif(not all(os.path.exists(os.path.join(DIR, f)) for f in ("ProPainter.pth", "raft-things.pth", "recurrent_flow_completion.pth")),
                                reason="ProPainter weights not staged under weights/propainter")


@pytest.mark.slow
@pytest.mark.skipif(os.environ.get("VSR_SLOW_TESTS") != "1", reason="the same run and the same check are part of tests/test_distributed_cpu.py (rank 0 of the "
                    "sharded-vs-single test) and tests/test_propainter_hybrid_cpu.py repeats it with the real kernels: set VSR_SLOW_TESTS=1 to run it alone")
def test_propainter_pipeline_on_cpu_runtime_equals_reference_frames():
    from fake_rt import FakeRuntime
    from make_golden_propainter import inputs
    from vsr_b200.propainter_inpaint import PropainterInpaint

    z = np.load(os.path.join(GOLDEN, "propainter_real.npz"))
    frames, mask = inputs()[:2]
    eng = PropainterInpaint("cuda:0", DIR, runtime=FakeRuntime())
    keep = [f.copy() for f in frames]
    out = np.stack(eng.inpaint(frames, mask))
    assert all(np.array_equal(a, b) for a, b in zip(frames, keep))
    d = np.abs(out.astype(np.int32) - z["comp"])
    assert out.shape == z["comp"].shape and d.max() <= 3 and (d > 0).mean() < 0.02, (int(d.max()), float((d > 0).mean()))
    hole = z["comp"] != np.stack(frames)
    assert hole.any() and np.abs(out[hole].astype(np.int32) - np.stack(frames)[hole]).mean() > 3        # the hole really was repainted


@pytest.mark.slow
@pytest.mark.skipif(os.environ.get("VSR_SLOW_TESTS") != "1", reason="about a minute on the CPU stand-in: set VSR_SLOW_TESTS=1 (last run: 1 grey level on 3e-5 of the pixels)")
def test_propainter_call_strips_on_cpu_runtime_equal_reference_frames():
    """`PropainterInpaint.__call__` (P1: strips with heights that are multiples of 8, first-frame mask, strip written back whole) against the
    unmodified reference's output for the 200x704 fixture (`call` in the golden file)."""
    from fake_rt import FakeRuntime
    from make_golden_propainter import inputs
    from vsr_b200.propainter_inpaint import PropainterInpaint

    z = np.load(os.path.join(GOLDEN, "propainter_real.npz"))
    big, big_mask = inputs()[2:]
    keep = [f.copy() for f in big]
    out = np.stack(PropainterInpaint("cuda:0", DIR, runtime=FakeRuntime())(big, big_mask))
    assert all(np.array_equal(a, b) for a, b in zip(big, keep))
    d = np.abs(out.astype(np.int32) - z["call"])
    assert out.shape == z["call"].shape and d.max() <= 3 and (d > 0).mean() < 0.02, (int(d.max()), float((d > 0).mean()))


def test_sub_ranges_follow_the_reference_chunk_arithmetic():
    """propainter_inpaint.py:254-272 / :284-304 restated: chunk [s_f, e_f) and the slice [pad_len_s, e_f - s_f - pad_len_e) kept from it."""
    from oracle import propainter_gen_oracle as G
    from vsr_b200 import propainter_tools as PT

    def reference(length, sub, pad_len):
        out = []
        for f in range(0, length, sub):
            s_f, e_f = max(0, f - pad_len), min(length, f + sub + pad_len)
            pad_len_s, pad_len_e = max(0, f) - s_f, e_f - min(length, f + sub)
            out.append((s_f, e_f, pad_len_s, e_f - s_f - pad_len_e))
        return out

    for length, sub, pad in ((13, 4, 5), (14, 4, 10), (250, 80, 5), (251, 100, 10), (79, 80, 5), (1, 1, 5)):
        want = reference(length, sub, pad)
        assert PT.sub_ranges(length, sub, pad) == want == G.sub_ranges(length, sub, pad)
        kept = [i for s, e, ks, ke in want for i in range(s + ks, s + ke)]
        assert kept == list(range(length))                     # the kept slices tile the sequence exactly once


@pytest.mark.slow
@pytest.mark.skipif(os.environ.get("VSR_SLOW_TESTS") != "1", reason="two minutes per variant on the CPU: set VSR_SLOW_TESTS=1 (last run: oracle and device pipeline "
                    "on the stand-in both within 1 grey level on 2e-5 of the pixels of the reference's frames)")
@pytest.mark.parametrize("which", ["oracle", "stand_in"])
def test_long_sequence_chunks_equal_reference_frames(which):
    """14 frames with sub_video_length = 4: overlapped chunks of flow completion and image propagation, capped reference frames
    (propainter_inpaint.py:251-324) against the UNMODIFIED reference (tests/golden/propainter_long.npz, tools/make_golden_propainter.py long)."""
    from make_golden_propainter import long_inputs

    z = np.load(os.path.join(GOLDEN, "propainter_long.npz"))["comp"]
    frames, mask, sub = long_inputs()
    if which == "oracle":
        from oracle import propainter_gen_oracle as G
        from oracle import raft_oracle as R
        from oracle import rfc_oracle as C

        w = {"raft": R.load_weights(os.path.join(DIR, "raft-things.pth")), "rfc": C.load_weights(os.path.join(DIR, "recurrent_flow_completion.pth")),
             "gen": G.load_weights(os.path.join(DIR, "ProPainter.pth"))}
        out = np.stack(G.inpaint(w, frames, mask, sub_video_length=sub))
    else:
        from fake_rt import FakeRuntime
        from vsr_b200.propainter_inpaint import PropainterInpaint

        out = np.stack(PropainterInpaint("cuda:0", DIR, sub_video_length=sub, runtime=FakeRuntime()).inpaint(frames, mask))
    d = np.abs(out.astype(np.int32) - z)
    assert d.max() <= 3 and (d > 0).mean() < 0.02, (int(d.max()), float((d > 0).mean()))
