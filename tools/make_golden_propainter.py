#!/usr/bin/env python
"""Generate tests/golden/propainter_real.npz by running the UNMODIFIED reference PropainterInpaint (/root/reference) on the
CPU with the reference's weights (ProPainter.pth = its 4 parts concatenated in fs_manifest.csv order, raft-things.pth,
recurrent_flow_completion.pth under weights/propainter/).  Build container only.

One dependency is replaced, not the reference: `torchvision.ops.deform_conv2d` (third party; its CPU kernel crashes on the
reference's shapes in this image) -> oracle/deform_conv.py, a pure-torch restatement checked against torchvision where
torchvision survives (tests/test_propainter_oracle.py).  Stored (fp16 where float) for a 128x192, 7-frame clip:

  gt_flows_f/b     RAFT_bi output            [1,T-1,2,H,W]
  pred_flows_f/b   RecurrentFlowCompleteNet + combine_flow (P4)
  prop_frames, prop_masks   InpaintGenerator.img_propagation (P5): updated frames [-1,1] and updated masks
  comp             PropainterInpaint.inpaint output frames (P6 + P7), uint8 BGR
  call             PropainterInpaint.__call__ output on the full frames (P1 + everything), uint8 BGR
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, sttn_oracle as O  # noqa: E402
from oracle.deform_conv import deform_conv2d  # noqa: E402


def inputs():
    # RAFT's 4-level correlation pyramid normalises sample coordinates by (H/64 - 1): inputs below 128 rows give 0/0 = NaN
    # flows in the reference itself (and black holes in its output), so the fixtures use >= 128 rows everywhere.
    H, W, T = 128, 192, 7
    frames = O.synthetic_clip(T, H, W, seed=23)
    mask = O.create_mask((H, W), [(40, 150, 84, 104)])
    big = O.synthetic_clip(5, 200, 704, seed=24)        # __call__: strip of int(704*3/16) = 132 -> 128 rows (multiple of 8)
    big_mask = O.create_mask((200, 704), [(120, 590, 150, 176)])
    return frames, mask, big, big_mask


def main():
    import torchvision

    torchvision.ops.deform_conv2d = deform_conv2d
    ref_import.install()
    from backend.inpaint.propainter_inpaint import PropainterInpaint

    torch.manual_seed(0)
    m = PropainterInpaint(torch.device("cpu"), os.path.join(ROOT, "weights", "propainter"), sub_video_length=80, use_fp16=False)
    taps = {}
    raft, rfc, gen = m.fix_raft, m.fix_flow_complete, m.model
    raft_fwd, combine, prop = raft.forward, rfc.combine_flow, gen.img_propagation

    def raft_hook(*a, **k):
        out = raft_fwd(*a, **k)
        taps.setdefault("gt_flows_f", out[0].clone())
        taps.setdefault("gt_flows_b", out[1].clone())
        return out

    def combine_hook(*a, **k):
        out = combine(*a, **k)
        taps.setdefault("pred_flows_f", out[0].clone())
        taps.setdefault("pred_flows_b", out[1].clone())
        return out

    def prop_hook(*a, **k):
        out = prop(*a, **k)
        taps.setdefault("prop_frames", out[0].clone())
        taps.setdefault("prop_masks", out[1].clone())
        return out

    raft.forward, rfc.combine_flow, gen.img_propagation = raft_hook, combine_hook, prop_hook
    frames, mask, big, big_mask = inputs()
    comp = m.inpaint([f.copy() for f in frames], mask)
    raft.forward, rfc.combine_flow, gen.img_propagation = raft_fwd, combine, prop
    call = m([f.copy() for f in big], big_mask)
    p = os.path.join(ROOT, "tests", "golden", "propainter_real.npz")
    np.savez_compressed(p, comp=np.stack(comp), call=np.stack(call),
                        **{k: v.numpy().astype(np.float16) for k, v in taps.items()})
    print(p, os.path.getsize(p), {k: tuple(v.shape) for k, v in taps.items()}, np.stack(comp).shape, np.stack(call).shape)


def long_inputs():
    """14 frames with sub_video_length = 4: the overlapped chunks of flow completion and image propagation (propainter_inpaint.py:251-312)
    and the capped reference frames of the window loop (:321-324) all take their long-sequence branches."""
    H, W, T = 128, 192, 14
    return O.synthetic_clip(T, H, W, seed=25), O.create_mask((H, W), [(40, 150, 84, 104)]), 4


def main_long():
    import torchvision

    torchvision.ops.deform_conv2d = deform_conv2d
    ref_import.install()
    from backend.inpaint.propainter_inpaint import PropainterInpaint

    torch.manual_seed(0)
    frames, mask, sub = long_inputs()
    m = PropainterInpaint(torch.device("cpu"), os.path.join(ROOT, "weights", "propainter"), sub_video_length=sub, use_fp16=False)
    comp = m.inpaint([f.copy() for f in frames], mask)
    p = os.path.join(ROOT, "tests", "golden", "propainter_long.npz")
    np.savez_compressed(p, comp=np.stack(comp))
    print(p, os.path.getsize(p), np.stack(comp).shape)


if __name__ == "__main__":
    main_long() if sys.argv[1:] == ["long"] else main()
