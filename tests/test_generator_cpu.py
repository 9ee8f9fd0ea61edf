"""CPU: the front half of the ProPainter generator's device graph (vsr_b200.propainter_generator: encoder with grouped convs, 1/4
flows and masks, learnable flow-guided feature propagation) on the fp32 stand-in of the runtime (kernel transcriptions) against
the stage taps of the oracle, which reproduces the reference's frames."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
from oracle import propainter_gen_oracle as G
from oracle import propainter_oracle as P

sys.path.insert(0, os.path.join(ROOT, "tools"))
PATH = os.path.join(ROOT, "weights", "propainter", "ProPainter.pth")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="ProPainter.pth not staged under weights/propainter")


def test_encoder_and_feature_propagation_on_cpu_runtime():
    from fake_rt import FakeRuntime
    from make_golden_propainter import inputs
    from vsr_b200.dbnet import _Tensor
    from vsr_b200.flow_propagation import propagate_images
    from vsr_b200.propainter_generator import Generator

    z = np.load(os.path.join(GOLDEN, "propainter_real.npz"))
    frames, mask = inputs()[:2]
    T, (H, W) = len(frames), frames[0].shape[:2]
    _, md = P.read_mask(mask, T)
    pf, pb = z["pred_flows_f"][0].astype(np.float32), z["pred_flows_b"][0].astype(np.float32)
    rt = FakeRuntime()
    gen = Generator(PATH, runtime=rt)
    x = _Tensor(rt.alloc(T * H * W * 8 * 2), 3, H, W, 8, n=T)
    rt.frames(frames, x)
    mask_dev, ff, fb = rt.upload_bytes(md[0]), rt.upload_bytes(pf), rt.upload_bytes(pb)
    state = propagate_images(rt, x, mask_dev, ff, fb)
    nb, refs = [0, 1, 2, 3], [6]           # 4 local frames + one reference frame (the 7-frame fixture's own windows have no references)
    ids = nb + refs
    enc, masks = gen.encode_and_propagate(state, mask_dev, ids, ff, fb, len(nb))
    # oracle taps for the same window
    xt = torch.from_numpy(np.stack([f[:, :, ::-1] for f in frames]).astype(np.float32) / 255).permute(0, 3, 1, 2)[None] * 2 - 1
    mt = torch.from_numpy(np.stack(md).astype(np.float32) / 255)[None, :, None]
    tf, tb = torch.from_numpy(pf)[None], torch.from_numpy(pb)[None]
    prop, upd = P.img_propagation(xt * (1 - mt), tf, tb, mt)
    updated = P.updated_frames(xt, mt, prop)
    taps = {}
    G.generator(G.load_weights(PATH), updated[:, ids], tf[:, nb[:-1]], tb[:, nb[:-1]], mt[:, ids], upd[:, ids], len(nb), taps)
    got = rt._v4(enc)[..., :128].transpose(0, 3, 1, 2)
    want = taps["enc_prop"][0].numpy()
    assert got.shape == want.shape
    scale = np.abs(want).max()
    assert np.abs(got[len(nb):] - want[len(nb):]).max() < 1e-4 * scale          # reference frames: encoder only
    assert np.abs(got[:len(nb)] - want[:len(nb)]).max() < 2e-3 * scale          # local frames: + deformable propagation
    # back half: soft split, 8 sparse-window transformer blocks, soft composition, decoder -> u8 predictions
    pred = gen.transform_and_decode(enc, len(nb), md[0], H, W)
    ref = G.generator(G.load_weights(PATH), updated[:, ids], tf[:, nb[:-1]], tb[:, nb[:-1]], mt[:, ids], upd[:, ids], len(nb))
    want8 = (((ref.view(-1, 3, H, W) + 1) / 2).permute(0, 2, 3, 1).numpy() * 255).astype(np.uint8)
    d = np.abs(pred.astype(np.int32) - want8)
    assert pred.shape == want8.shape and d.max() <= 2 and (d > 0).mean() < 0.02, (int(d.max()), float((d > 0).mean()))
