"""CPU: the RAFT device graph (vsr_b200.raft_flow.RaftFlow: folded batch-norm encoder, instance-norm encoder, correlation GEMM +
pyramid + lookup, the GRU state tensors, graph replay of the iterations, convex up-sampling) driven on the fp32 stand-in of the
runtime against the oracle, which is pinned to the reference's RAFT_bi flows."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import raft_oracle as R
from oracle import sttn_oracle as O

PATH = os.path.join(ROOT, "weights", "propainter", "raft-things.pth")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="raft-things.pth not staged under weights/propainter")


def test_raft_graph_on_cpu_runtime():
    from fake_rt import FakeRuntime
    from vsr_b200.raft_flow import RaftFlow

    frames = O.synthetic_clip(3, 128, 192, seed=23)
    rt = FakeRuntime()
    eng = RaftFlow(PATH, runtime=rt)
    ff, fb = eng(frames, iters=6)
    x = torch.from_numpy(np.stack([f[:, :, ::-1] for f in frames]).astype(np.float32) / 255).permute(0, 3, 1, 2)[None] * 2 - 1
    wf, wb = R.raft_bi(R.load_weights(PATH), x, iters=6)
    assert ff.shape == fb.shape == (2, 2, 128, 192)
    assert np.abs(ff - wf[0].numpy()).max() < 2e-3 and np.abs(fb - wb[0].numpy()).max() < 2e-3
    assert np.abs(ff).max() > 0.5                       # the synthetic clip really moves
    n0 = rt.launch_count
    ff2, _ = eng(frames, iters=6)                       # second call: recorded iteration graphs
    assert np.array_equal(ff2, ff) and rt.launch_count > n0
