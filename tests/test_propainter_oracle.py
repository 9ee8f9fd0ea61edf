"""CPU: the parts of the ProPainter oracle that exist (oracle/propainter_oracle.py: P1, P2, P5, P7) against the golden
taps of the unmodified reference (tests/golden/propainter_real.npz), and the pure-torch deform_conv2d against torchvision."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
from oracle import propainter_oracle as P
from oracle import sttn_oracle as O

sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_deform_conv_matches_torchvision():
    tv = pytest.importorskip("torchvision")
    from oracle.deform_conv import deform_conv2d

    torch.manual_seed(0)
    for (cin, cout, G, H, W, s, p, d) in [(4, 6, 1, 6, 7, 1, 1, 1), (8, 8, 2, 9, 10, 1, 1, 1), (16, 8, 4, 12, 9, 2, 1, 1), (8, 4, 8, 7, 7, 1, 2, 2)]:
        x, w = torch.randn(2, cin, H, W), torch.randn(cout, cin, 3, 3)
        Ho, Wo = (H + 2 * p - 2 * d - 1) // s + 1, (W + 2 * p - 2 * d - 1) // s + 1
        off, m, b = torch.randn(2, 2 * G * 9, Ho, Wo) * 2, torch.rand(2, G * 9, Ho, Wo), torch.randn(cout)
        ref = tv.ops.deform_conv2d(x, off, w, b, s, p, d, m)
        assert (deform_conv2d(x, off, w, b, s, p, d, m) - ref).abs().max() < 1e-4


def test_read_mask_equals_scipy():
    sp = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(3)
    for _ in range(5):
        m = (rng.random((40, 60)) > 0.97).astype(np.uint8) * 255
        m[0, 0] = m[-1, -1] = 255
        for it in (1, 4, 8):
            assert np.array_equal(P.binary_dilation_cross(m, it), sp.binary_dilation(m, iterations=it))
    fm, md = P.read_mask(m[:, :, None], 3)
    assert len(fm) == len(md) == 3 and fm[0].dtype == np.uint8 and set(np.unique(md[0])) <= {0, 255}
    assert np.array_equal(md[0] > 0, sp.binary_dilation(m, iterations=4))


def test_window_schedule_and_refs():
    s = P.window_schedule(23)
    assert [w[0][0] for w in s] == [0, 0, 5, 10, 15] and s[2][0] == list(range(5, 16)) and s[2][1] == [0, 20]
    assert P.get_ref_index(40, list(range(35, 46)), 200, 10, 8) == [0, 10, 20, 30, 50, 60, 70]
    long = P.window_schedule(120, 80)
    assert all(len(r) <= 9 for _, r in long)


def test_image_propagation_and_composite_against_reference_taps():
    from make_golden_propainter import inputs

    z = np.load(os.path.join(GOLDEN, "propainter_real.npz"))
    frames, mask, big, big_mask = inputs()
    T, H, W = len(frames), frames[0].shape[0], frames[0].shape[1]
    _, md = P.read_mask(mask, T)
    masks = torch.from_numpy(np.stack(md).astype(np.float32) / 255)[None, :, None]                    # to_tensors(): /255
    x = torch.from_numpy(np.stack([f[:, :, ::-1] for f in frames]).astype(np.float32) / 255).permute(0, 3, 1, 2)[None] * 2 - 1
    ff, fb = (torch.from_numpy(z[k].astype(np.float32)) for k in ("pred_flows_f", "pred_flows_b"))
    prop, upd = P.img_propagation(x * (1 - masks), ff, fb, masks)
    # the stored flows are fp16 roundings of the reference's: nearest warps flip on a handful of pixels at most
    assert (upd != torch.from_numpy(z["prop_masks"].astype(np.float32))).float().mean() < 2e-3
    d = (prop - torch.from_numpy(z["prop_frames"].astype(np.float32))).abs()
    assert (d > 2e-3).float().mean() < 2e-3
    # pixels outside the dilated mask are never touched; the hole shrinks where neighbours were known
    assert torch.equal(prop * (1 - masks), x * (1 - masks)) and upd.sum() < masks.sum()
    # P1: the strip of __call__ and what lies outside it
    (y0, y1, x0, x1), = P.strip_areas(704, 200, big_mask)
    assert (y1 - y0) % 8 == 0 and (y1 - y0) == 128 and (x0, x1) == (0, 704)   # int(704*3/16) = 132 rounded to a multiple of 8
    for o, f in zip(z["call"], big):
        assert np.array_equal(o[:y0], f[:y0]) and np.array_equal(o[y1:], f[y1:])
    # P7: composite restated on the reference's own final frames: known pixels of `comp` equal the input
    keep = np.stack(md) == 0
    assert np.array_equal(z["comp"][keep], np.stack(frames)[keep])
    c = P.composite([None] * 2, np.full((2, 4, 4, 3), 200.7, np.float32), np.ones((2, 4, 4, 1), np.uint8), [np.zeros((4, 4, 3), np.uint8)] * 2, [0, 1])
    c = P.composite(c, np.full((1, 4, 4, 3), 101.9, np.float32), np.ones((1, 4, 4, 1), np.uint8), [np.zeros((4, 4, 3), np.uint8)] * 2, [1])
    assert c[0][0, 0, 0] == 200 and c[1][0, 0, 0] == 150 and c[1].dtype == np.uint8


def test_raft_oracle_matches_reference_flows():
    """P3: the RAFT restatement against RAFT_bi of the unmodified reference (golden gt_flows, stored as fp16)."""
    path = os.path.join(ROOT, "weights", "propainter", "raft-things.pth")
    if not os.path.exists(path):
        pytest.skip("raft-things.pth not staged under weights/propainter")
    from make_golden_propainter import inputs
    from oracle import raft_oracle as R

    z = np.load(os.path.join(GOLDEN, "propainter_real.npz"))
    frames = inputs()[0][:3]
    x = torch.from_numpy(np.stack([f[:, :, ::-1] for f in frames]).astype(np.float32) / 255).permute(0, 3, 1, 2)[None] * 2 - 1
    ff, fb = R.raft_bi(R.load_weights(path), x, iters=20)
    for got, key in ((ff, "gt_flows_f"), (fb, "gt_flows_b")):
        want = torch.from_numpy(z[key][:, :2].astype(np.float32))
        assert (got - want).abs().max() < 4e-3          # fp16 storage of flows of magnitude ~3 px: half an ulp is 1e-3


def test_flow_completion_oracle_matches_reference():
    """P4: RecurrentFlowCompleteNet restated (with the pure-torch deformable conv) against the reference's completed flows."""
    path = os.path.join(ROOT, "weights", "propainter", "recurrent_flow_completion.pth")
    if not os.path.exists(path):
        pytest.skip("recurrent_flow_completion.pth not staged under weights/propainter")
    from make_golden_propainter import inputs
    from oracle import rfc_oracle as C

    z = np.load(os.path.join(GOLDEN, "propainter_real.npz"))
    frames, mask = inputs()[:2]
    fm, _ = P.read_mask(mask, len(frames))
    masks = torch.from_numpy(np.stack(fm).astype(np.float32) / 255)[None, :, None]
    gf, gb = (torch.from_numpy(z[k].astype(np.float32)) for k in ("gt_flows_f", "gt_flows_b"))
    pf, pb = C.complete_bidirectional(C.load_weights(path), gf, gb, masks)
    for got, key in ((pf, "pred_flows_f"), (pb, "pred_flows_b")):
        want = torch.from_numpy(z[key].astype(np.float32))
        assert (got - want).abs().max() < 2e-2, float((got - want).abs().max())     # inputs and pins are fp16 roundings (~2e-3 each)
    hole = masks[:, :-1].expand_as(pf) > 0
    assert (pf[hole] - gf[hole]).abs().mean() > 1e-3 and torch.equal(pf[~hole], gf[~hole])


def _all_weights():
    d = os.path.join(ROOT, "weights", "propainter")
    if not all(os.path.exists(os.path.join(d, f)) for f in ("ProPainter.pth", "raft-things.pth", "recurrent_flow_completion.pth")):
        pytest.skip("ProPainter weights not staged under weights/propainter")
    from oracle import propainter_gen_oracle as G
    from oracle import raft_oracle as R
    from oracle import rfc_oracle as C

    return G, dict(raft=R.load_weights(os.path.join(d, "raft-things.pth")), rfc=C.load_weights(os.path.join(d, "recurrent_flow_completion.pth")),
                   gen=G.load_weights(os.path.join(d, "ProPainter.pth")))


@pytest.mark.slow
def test_whole_inpaint_chain_matches_reference_frames():
    """P2 -> P3 -> P4 -> P5 -> P6 -> P7 restated end to end against `PropainterInpaint.inpaint` of the unmodified reference."""
    from make_golden_propainter import inputs

    G, weights = _all_weights()
    z = np.load(os.path.join(GOLDEN, "propainter_real.npz"))
    frames, mask = inputs()[:2]
    taps = {}
    out = np.stack(G.inpaint(weights, frames, mask, taps=taps))
    for k in ("gt_flows_f", "pred_flows_f", "prop_frames"):
        assert (taps[k] - torch.from_numpy(z[k].astype(np.float32))).abs().max() < 2e-2
    d = np.abs(out.astype(np.int32) - z["comp"])
    assert d.max() <= 2 and (d > 0).mean() < 5e-3, (int(d.max()), float((d > 0).mean()))     # fp32 re-association before the u8 truncation


@pytest.mark.slow
def test_call_strips_match_reference_frames():
    """P1 + the chain: `PropainterInpaint.__call__` (strip of 128 rows at 704 px width)."""
    from make_golden_propainter import inputs

    G, weights = _all_weights()
    z = np.load(os.path.join(GOLDEN, "propainter_real.npz"))
    _, _, big, big_mask = inputs()
    keep = [f.copy() for f in big]
    out = np.stack(G.propainter_call(weights, big, big_mask))
    assert all(np.array_equal(a, b) for a, b in zip(big, keep))
    d = np.abs(out.astype(np.int32) - z["call"])
    assert d.max() <= 2 and (d > 0).mean() < 5e-3, (int(d.max()), float((d > 0).mean()))
