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
