"""CPU oracle for the ProPainter path — TEST INFRASTRUCTURE ONLY.  This file: the host / integer
rows and the image propagation; RAFT is oracle/raft_oracle.py (P3), flow completion oracle/rfc_oracle.py (P4), the generator
and the whole `inpaint` / `__call__` chain oracle/propainter_gen_oracle.py (P6).

Restated and pinned against the unmodified reference (tests/golden/propainter_real.npz, tools/make_golden_propainter.py):
  P1  strips for PropainterInpaint.__call__              backend/inpaint/propainter_inpaint.py:363-418
  P2  read_mask (binary dilation, iterations=4)          propainter_inpaint.py:32-77
  P5  InpaintGenerator.img_propagation (learnable=False) video/model/propainter.py:24-33,107-193,316-319;
      flow_warp                                          video/model/modules/flow_loss_utils.py:6-45
  P7  window loop, get_ref_index, composite + blend      propainter_inpaint.py:120-135,318-361
Parity: PINNED.  The chained oracle reproduces the reference's final frames of `inpaint` and `__call__` up to 1 grey level on
2e-5 of the pixels (fp32 re-association before the u8 truncation).  `oracle/deform_conv.py` is the pure-torch `deform_conv2d`
P4 and P6 need (torchvision's CPU kernel crashes on the reference's shapes in this image).
"""
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

NEIGHBOR_LENGTH, REF_STRIDE, MASK_DILATION, RAFT_ITERS = 10, 10, 4, 20   # propainter_inpaint.py:148-158


# ------------------------------------------------------------------------------------------------ P1 / P2: host integer path
def strip_areas(W: int, H: int, mask: np.ndarray):
    """__call__ :373-374: strips of int(W*3/16) rows, heights rounded to multiples of 8 (`multiple=8`)."""
    from oracle import sttn_oracle as O

    return O.get_inpaint_area_by_mask(W, H, int(W * 3 / 16), mask, multiple=8)


def binary_dilation_cross(m: np.ndarray, iterations: int) -> np.ndarray:
    """scipy.ndimage.binary_dilation(m, iterations=n) with its default structuring element (3x3 cross, border value 0)."""
    a = m.astype(bool)
    for _ in range(iterations):
        b = a.copy()
        b[1:] |= a[:-1]
        b[:-1] |= a[1:]
        b[:, 1:] |= a[:, :-1]
        b[:, :-1] |= a[:, 1:]
        a = b
    return a


def read_mask(mask: np.ndarray, length: int, flow_mask_dilates: int = MASK_DILATION, mask_dilates: int = MASK_DILATION):
    """read_mask for the ndarray input the pipeline uses (:36-43,55-75): -> (flow_masks, masks_dilated), each `length` u8
    arrays in {0, 255}.  With dilates == 0 the reference thresholds at 0.1 instead (binary_mask)."""
    m = mask.squeeze(2) if mask.ndim == 3 and mask.shape[2] == 1 else mask
    if m.ndim == 3:
        raise ValueError("colour masks are converted with cv2 in the reference; pass a single-channel mask")

    def one(it):
        return (binary_dilation_cross(m, it) if it > 0 else (m > 0.1)).astype(np.uint8) * 255

    return [one(flow_mask_dilates)] * length, [one(mask_dilates)] * length


# ------------------------------------------------------------------------------------------------ P5: image propagation
def flow_warp(x: torch.Tensor, flow: torch.Tensor, interpolation: str = "bilinear") -> torch.Tensor:
    """flow_loss_utils.py:6-45: x [n,c,h,w], flow [n,h,w,2] = (dx, dy) in pixels; zeros outside, align_corners=True."""
    _, _, h, w = x.shape
    gy, gx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    g = torch.stack((gx, gy), 2).type_as(x) + flow
    g = torch.stack((2.0 * g[..., 0] / max(w - 1, 1) - 1.0, 2.0 * g[..., 1] / max(h - 1, 1) - 1.0), 3)
    return F.grid_sample(x, g, mode=interpolation, padding_mode="zeros", align_corners=True)


def fb_consistency(flow_fw: torch.Tensor, flow_bw: torch.Tensor, alpha1=0.01, alpha2=0.5) -> torch.Tensor:
    """propainter.py:24-33: 1 where the forward flow and the backward flow sampled at its target cancel."""
    bw = flow_warp(flow_bw, flow_fw.permute(0, 2, 3, 1))
    sq = lambda t: torch.sum(torch.square(t), 1, keepdim=True)  # noqa: E731
    return (sq(flow_fw + bw) < alpha1 * (sq(flow_fw) + sq(bw)) + alpha2).to(flow_fw)


def _binary(t: torch.Tensor) -> torch.Tensor:
    return (t > 0.1).to(t)          # BidirectionalPropagation.binary_mask :102-106


def img_propagation(masked_frames, flows_f, flows_b, masks, interpolation="nearest"):
    """BidirectionalPropagation(learnable=False).forward (:107-193): backward pass over t = T-1..0 with the forward flows,
    then a forward pass over its output with the backward flows; a masked pixel takes the warped neighbour when the flows
    are consistent and the neighbour is known there.  Returns (prop_frames [b,t,c,h,w], updated_masks [b,t,1,h,w])."""
    b, t, c, h, w = masked_frames.shape
    feats = [masked_frames[:, i] for i in range(t)]
    msks = [masks[:, i] for i in range(t)]
    for direction in ("backward", "forward"):
        order = list(range(t - 1, -1, -1)) if direction == "backward" else list(range(t))
        out_f, out_m = [None] * t, [None] * t
        prev_f = prev_m = None
        for n, idx in enumerate(order):
            cur_f, cur_m = feats[idx], msks[idx]
            if n == 0:
                nf, nm = cur_f, cur_m
            else:
                fi = idx if direction == "backward" else idx - 1          # flow between frames fi and fi+1
                prop = (flows_f if direction == "backward" else flows_b)[:, fi]
                check = (flows_b if direction == "backward" else flows_f)[:, fi]
                valid = fb_consistency(prop, check)
                warped = flow_warp(prev_f, prop.permute(0, 2, 3, 1), interpolation)
                hole_there = _binary(flow_warp(prev_m, prop.permute(0, 2, 3, 1)))
                use = _binary(cur_m * valid * (1 - hole_there))
                nf = use * warped + (1 - use) * cur_f
                nm = _binary(cur_m * (1 - valid * (1 - hole_there)))
            out_f[idx], out_m[idx] = nf, nm
            prev_f, prev_m = nf, nm
        feats, msks = out_f, out_m
    return torch.stack(feats, 1), torch.stack(msks, 1)


def updated_frames(frames, masks_dilated, prop_frames):
    """propainter_inpaint.py:311: known pixels from the input, masked pixels from the propagation."""
    return frames * (1 - masks_dilated) + prop_frames * masks_dilated


# ------------------------------------------------------------------------------------------------ P7: window loop + composite
def get_ref_index(mid: int, neighbor_ids: Sequence[int], length: int, ref_stride: int = REF_STRIDE, ref_num: int = -1) -> List[int]:
    """propainter_inpaint.py:120-135."""
    if ref_num == -1:
        return [i for i in range(0, length, ref_stride) if i not in neighbor_ids]
    out: List[int] = []
    lo, hi = max(0, mid - ref_stride * (ref_num // 2)), min(length, mid + ref_stride * (ref_num // 2))
    for i in range(lo, hi, ref_stride):
        if i not in neighbor_ids:
            if len(out) > ref_num:
                break
            out.append(i)
    return out


def window_schedule(video_length: int, sub_video_length: int = 80) -> List[Tuple[List[int], List[int]]]:
    """:318-333: (neighbor_ids, ref_ids) per window; windows start every neighbor_length // 2 = 5 frames."""
    stride = NEIGHBOR_LENGTH // 2
    ref_num = sub_video_length // REF_STRIDE if video_length > sub_video_length else -1
    out = []
    for f in range(0, video_length, stride):
        nb = list(range(max(0, f - stride), min(video_length, f + stride + 1)))
        out.append((nb, get_ref_index(f, nb, video_length, REF_STRIDE, ref_num)))
    return out


def composite(comp_frames: List, pred_img: np.ndarray, binary_masks: np.ndarray, ori_frames: Sequence[np.ndarray], neighbor_ids: Sequence[int]):
    """:344-357: pred_img [n,h,w,3] float in 0..255 (RGB), binary_masks [n,h,w,1] u8 in {0,1}: u8 truncation, mask composite,
    0.5/0.5 blend with the previous visit, re-quantised to u8 EVERY time."""
    for i, idx in enumerate(neighbor_ids):
        img = np.array(pred_img[i]).astype(np.uint8) * binary_masks[i] + ori_frames[idx] * (1 - binary_masks[i])
        if comp_frames[idx] is None:
            comp_frames[idx] = img
        else:
            comp_frames[idx] = comp_frames[idx].astype(np.float32) * 0.5 + img.astype(np.float32) * 0.5
        comp_frames[idx] = comp_frames[idx].astype(np.uint8)
    return comp_frames
