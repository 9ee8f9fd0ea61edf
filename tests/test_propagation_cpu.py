"""CPU: the device graph of ProPainter's image propagation (vsr_b200.flow_propagation) on the fp32 stand-in of
the runtime against the oracle and the reference's own taps (tests/golden/propainter_real.npz)."""
import os
import sys

import numpy as np
import torch

from conftest import GOLDEN, ROOT
from oracle import propainter_oracle as P

sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_image_propagation_graph_on_cpu_runtime():
    from fake_rt import FakeRuntime
    from make_golden_propainter import inputs
    from vsr_b200.flow_propagation import propagate_images_host

    z = np.load(os.path.join(GOLDEN, "propainter_real.npz"))
    frames, mask = inputs()[:2]
    _, md = P.read_mask(mask, len(frames))
    ff, fb = z["pred_flows_f"][0].astype(np.float32), z["pred_flows_b"][0].astype(np.float32)
    upd, um = propagate_images_host(FakeRuntime(), frames, md[0], ff, fb)
    x = torch.from_numpy(np.stack([f[:, :, ::-1] for f in frames]).astype(np.float32) / 255).permute(0, 3, 1, 2)[None] * 2 - 1
    masks = torch.from_numpy(np.stack(md).astype(np.float32) / 255)[None, :, None]
    prop, want_m = P.img_propagation(x * (1 - masks), torch.from_numpy(ff)[None], torch.from_numpy(fb)[None], masks)
    want = P.updated_frames(x, masks, prop)[0].numpy()
    assert np.abs(upd - want).max() < 1e-5 and np.array_equal(um, want_m[0].numpy())
    # and against the reference's own taps (stored as fp16; the flows fed in are fp16 roundings too)
    assert (np.abs(um - z["prop_masks"][0].astype(np.float32)) > 0).mean() < 2e-3
