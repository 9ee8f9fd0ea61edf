"""GPU: RAFT on the device runtime against the oracle.  GATED: the device path was written after round 1's
GPU budget was spent and has not run on a B200 yet; set VSR_RUN_UNVALIDATED=1 to run it (the first thing to do next round).
Tolerance to establish then (fp16 features, fp32 flow state): mean end-point error <= 0.05 px, max <= 0.5 px on the fixtures."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import raft_oracle as R
from oracle import sttn_oracle as O

PATH = os.path.join(ROOT, "weights", "propainter", "raft-things.pth")
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("VSR_RUN_UNVALIDATED") != "1", reason="RAFT device path not yet validated on a B200 (DESIGN.md §7)"),
              pytest.mark.skipif(not os.path.exists(PATH), reason="raft-things.pth not staged under weights/propainter")]


def test_raft_flows_vs_oracle(capi):
    from vsr_b200.raft_flow import RaftFlow

    frames = O.synthetic_clip(3, 128, 192, seed=23)
    ff, fb = RaftFlow(PATH, "cuda:0")(frames)
    x = torch.from_numpy(np.stack([f[:, :, ::-1] for f in frames]).astype(np.float32) / 255).permute(0, 3, 1, 2)[None] * 2 - 1
    wf, wb = R.raft_bi(R.load_weights(PATH), x)
    for got, want in ((ff, wf[0].numpy()), (fb, wb[0].numpy())):
        epe = np.sqrt(((got - want) ** 2).sum(1))
        assert np.isfinite(got).all() and epe.mean() <= 0.05 and epe.max() <= 0.5, (float(epe.mean()), float(epe.max()))
