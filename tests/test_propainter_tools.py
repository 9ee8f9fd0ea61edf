"""CPU: the host rows of the ProPainter path that exist in the product (vsr_b200.propainter_tools: P1, P2, P7) against scipy,
the oracle, and the golden frames of the unmodified reference."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import propainter_oracle as P

sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_strips_masks_schedule_equal_oracle_and_scipy():
    sp = pytest.importorskip("scipy.ndimage")
    from vsr_b200 import create_mask
    from vsr_b200 import propainter_tools as T

    rng = np.random.default_rng(8)
    for _ in range(20):
        H, W = int(rng.integers(200, 500)), int(rng.integers(300, 900))
        x0, y0 = int(rng.integers(0, W - 60)), int(rng.integers(0, H - 30))
        mask = create_mask((H, W), [(x0, x0 + int(rng.integers(20, 59)), y0, y0 + int(rng.integers(8, 29)))])
        assert T.strip_areas(W, H, mask) == P.strip_areas(W, H, mask)
        assert all((a[1] - a[0]) % 8 == 0 for a in T.strip_areas(W, H, mask))
        fm, md = T.read_mask(mask[:, :, None], 3)
        assert np.array_equal(md[0] > 0, sp.binary_dilation(mask, iterations=4)) and np.array_equal(fm[0], P.read_mask(mask, 3)[0][0])
    for n, sub in ((7, 80), (23, 80), (80, 80), (120, 80), (301, 80)):
        assert T.window_schedule(n, sub) == P.window_schedule(n, sub)
    assert T.get_ref_index(40, list(range(35, 46)), 200, 10, 8) == P.get_ref_index(40, list(range(35, 46)), 200, 10, 8)


def test_composite_reproduces_reference_frames_from_reference_predictions():
    """P7 on real data: feeding the generator outputs of the ORACLE chain (which equals the reference's) through the product's
    window schedule + composite gives the reference's final frames."""
    d = os.path.join(ROOT, "weights", "propainter")
    if not all(os.path.exists(os.path.join(d, f)) for f in ("ProPainter.pth", "raft-things.pth", "recurrent_flow_completion.pth")):
        pytest.skip("ProPainter weights not staged under weights/propainter")
    import torch
    from make_golden_propainter import inputs
    from oracle import propainter_gen_oracle as G
    from oracle import raft_oracle as R
    from oracle import rfc_oracle as C
    from vsr_b200 import propainter_tools as T

    w = dict(raft=R.load_weights(os.path.join(d, "raft-things.pth")), rfc=C.load_weights(os.path.join(d, "recurrent_flow_completion.pth")),
             gen=G.load_weights(os.path.join(d, "ProPainter.pth")))
    z = np.load(os.path.join(GOLDEN, "propainter_real.npz"))
    frames, mask = inputs()[:2]
    n, (H, W) = len(frames), frames[0].shape[:2]
    fm, md = T.read_mask(mask, n)
    x = torch.from_numpy(np.stack([f[:, :, ::-1] for f in frames]).astype(np.float32) / 255).permute(0, 3, 1, 2)[None] * 2 - 1
    masks = torch.from_numpy(np.stack(md).astype(np.float32) / 255)[None, :, None]
    pf, pb = (torch.from_numpy(z[k].astype(np.float32)) for k in ("pred_flows_f", "pred_flows_b"))
    prop, upd = P.img_propagation(x * (1 - masks), pf, pb, masks)
    updated = P.updated_frames(x, masks, prop)
    comp = [None] * n
    binary = masks[0].permute(0, 2, 3, 1).numpy().astype(np.uint8)
    rgb = [np.ascontiguousarray(f[:, :, ::-1]) for f in frames]
    for nb, refs in T.window_schedule(n):
        ids = nb + refs
        pred = G.generator(w["gen"], updated[:, ids], pf[:, nb[:-1]], pb[:, nb[:-1]], masks[:, ids], upd[:, ids], len(nb))
        pred = ((pred.view(-1, 3, H, W) + 1) / 2).permute(0, 2, 3, 1).numpy() * 255
        T.composite(comp, pred, binary[nb], rgb, nb)
    out = np.stack([c[:, :, ::-1] for c in comp])
    d_ = np.abs(out.astype(np.int32) - z["comp"])
    assert d_.max() <= 2 and (d_ > 0).mean() < 5e-3
