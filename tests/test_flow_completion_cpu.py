"""CPU: the device graph of ProPainter's flow completion (vsr_b200.flow_completion) on the fp32 stand-in of the
runtime (whose operators are transcriptions of the csrc/pp_ops.cuh kernels) against the oracle and the reference's taps."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
from oracle import propainter_oracle as P
from oracle import rfc_oracle as C

sys.path.insert(0, os.path.join(ROOT, "tools"))
PATH = os.path.join(ROOT, "weights", "propainter", "recurrent_flow_completion.pth")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="recurrent_flow_completion.pth not staged under weights/propainter")


def test_flow_completion_graph_on_cpu_runtime():
    from fake_rt import FakeRuntime
    from make_golden_propainter import inputs
    from vsr_b200.flow_completion import FlowCompletion

    z = np.load(os.path.join(GOLDEN, "propainter_real.npz"))
    frames, mask = inputs()[:2]
    fm, _ = P.read_mask(mask, len(frames))
    gf, gb = z["gt_flows_f"][0, :4].astype(np.float32), z["gt_flows_b"][0, :4].astype(np.float32)     # 4 flows = 5 frames
    pf, pb = FlowCompletion(PATH, runtime=FakeRuntime()).complete_host(gf, gb, fm[0])
    masks = torch.from_numpy(np.stack(fm[:5]).astype(np.float32) / 255)[None, :, None]
    wf, wb = C.complete_bidirectional(C.load_weights(PATH), torch.from_numpy(gf)[None], torch.from_numpy(gb)[None], masks)
    assert np.abs(pf - wf[0].numpy()).max() < 2e-3 and np.abs(pb - wb[0].numpy()).max() < 2e-3
    hole = fm[0] > 0
    assert np.abs(pf[:, :, hole] - gf[:, :, hole]).mean() > 1e-3 and np.array_equal(pf[:, :, ~hole], gf[:, :, ~hole])


def test_second_call_reuses_buffers():
    from fake_rt import FakeRuntime
    from vsr_b200.flow_completion import FlowCompletion

    rng = np.random.default_rng(0)
    f = rng.standard_normal((3, 2, 64, 64)).astype(np.float32)
    mask = np.zeros((64, 64), np.uint8)
    mask[20:40, 10:50] = 255
    rt = FakeRuntime()
    eng = FlowCompletion(PATH, runtime=rt)
    a = eng.complete_host(f, -f, mask)
    n = len(rt.bufs)
    b = eng.complete_host(f, -f, mask)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert len(rt.bufs) - n <= 5            # only the uploaded inputs and the two result buffers are new
