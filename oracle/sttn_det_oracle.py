"""ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE) — CPU restatement of the reference STTN-det path
(SURVEY.md §8a rows D1-D3).  Same rules as oracle/sttn_oracle.py: imported only by tests/, smoke() and
bench.py's CPU legs.  Parity: PINNED against the unmodified reference by tools/make_golden.py ->
tests/golden/sttn_det_*.npz and tests/test_oracle_golden.py.

The network is the STTN-auto one with a different geometry (432x240 input, 108x60 feature map, patches
(108,60) (36,20) (18,10) (9,5)); the attention mask of network_sttn.py:149 is a no-op (`masked_fill`
result discarded), so `infer` is oracle.sttn_oracle.infer with this patch set."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import sttn_oracle as O

MODEL_W, MODEL_H = 432, 240  # backend/inpaint/sttn_det_inpaint.py:33
PATCHSIZE = [(108, 60), (36, 20), (18, 10), (9, 5)]  # backend/inpaint/sttn/network_sttn.py:70


def split_height(H: int, W: int) -> int:
    """sttn_det_inpaint.py:48-51."""
    return int(H * 5 / 9) if H > W else int(W * 5 / 18)


def inpaint_strip(w, frames_bgr: Sequence[np.ndarray], masks: Sequence[np.ndarray], stride: int = O.NEIGHBOR_STRIDE,
                  ref_length: int = O.REF_LENGTH) -> List[np.ndarray]:
    """STTNDetInpaint.inpaint, sttn_det_inpaint.py:124-174.  frames: T x [240,432,3] u8 BGR; masks: T x
    [240,432] u8 (the resized 0..255 mask).  Returns comps in RGB (uint8 single visit / float32 blended)."""
    T = len(frames_bgr)
    binary = [(np.asarray(m) > 0.5).astype(np.uint8)[:, :, None] for m in masks]  # :132  (any non-zero)
    # masks_tensor: Stack/ToTorchFormatTensor divide by 255, then > 0.5  => m >= 128   :134
    mt = torch.from_numpy(np.stack([(np.asarray(m).astype(np.float32) / 255.0 > 0.5) for m in masks]).astype(np.float32))[:, None]
    rgb = [f[:, :, ::-1] for f in frames_bgr]  # Stack() turned the caller's list into RGB images (utils/sttn_utils.py:71-75)
    with torch.no_grad():
        x = O.frames_to_tensor(frames_bgr)
        feats = O.encoder(w, x * (1 - mt))  # :143
        comps: List = [None] * T
        for nb, refs in O.window_schedule(T, stride, ref_length):
            pred = O.infer(w, feats[nb + refs], PATCHSIZE)
            img = O.quantise(O.decoder(w, pred[:len(nb)]))
            for i, idx in enumerate(nb):
                c = img[i] * binary[idx] + rgb[idx] * (1 - binary[idx])  # :168
                if comps[idx] is None:
                    comps[idx] = c
                else:
                    comps[idx] = comps[idx].astype(np.float32) * 0.5 + c.astype(np.float32) * 0.5
    return comps


def det_call(w, input_frames: Sequence[np.ndarray], input_mask: np.ndarray) -> List[np.ndarray]:
    """STTNDetInpaint.__call__, sttn_det_inpaint.py:38-99: mask is NOT thresholded, frame and mask strips are
    resized to 432x240, and the whole strip is replaced by the up-scaled comp (:93)."""
    mask = np.asarray(input_mask)
    H, W = mask.shape[:2]
    sh = split_height(H, W)
    areas = O.get_inpaint_area_by_mask(W, H, sh, mask)
    frames = [f.copy() for f in input_frames]
    if not areas:
        return frames
    comps = {}
    for k, (y0, y1, _, _) in enumerate(areas):
        scaled = [O.cv2_resize_linear_u8(np.ascontiguousarray(f[y0:y1]), MODEL_W, MODEL_H) for f in frames]
        m = O.cv2_resize_linear_u8(np.ascontiguousarray(mask[y0:y1]), MODEL_W, MODEL_H)
        comps[k] = inpaint_strip(w, scaled, [m] * len(frames))
    for j, frame in enumerate(frames):
        for k, (y0, y1, _, _) in enumerate(areas):
            frame[y0:y1] = O.upscale_comp(comps[k][j], W, sh)
    return frames
