"""Host-side rows of the ProPainter path — integer / index logic, bit-exact against the reference
(tests/test_propainter_tools.py runs them against scipy, the oracle and the golden frames of the unmodified reference).
The device rows (P3 RAFT, P4 flow completion, P5 image propagation, P6 generator) are not built yet: there is no
`PropainterInpaint` class in this package until they are; nothing here falls back to the CPU models."""
from typing import List, Sequence, Tuple

import numpy as np

from .inpaint_tools import get_inpaint_area_by_mask

NEIGHBOR_LENGTH = 10    # propainter_inpaint.py:148
REF_STRIDE = 10         # :152
MASK_DILATION = 4       # :150


def strip_areas(W: int, H: int, mask: np.ndarray):
    """P1 — PropainterInpaint.__call__ (propainter_inpaint.py:372-374): strips of int(W*3/16) rows around the mask rows, heights
    rounded to multiples of 8; (ymin, ymax, xmin, xmax) per strip.  Only the first frame's mask is used (:400)."""
    return get_inpaint_area_by_mask(W, H, int(W * 3 / 16), mask, multiple=8)


def binary_dilation_cross(mask: np.ndarray, iterations: int) -> np.ndarray:
    """scipy.ndimage.binary_dilation(mask, iterations=n) with the default 3x3 cross and border value 0 (what read_mask calls)."""
    a = np.asarray(mask).astype(bool)
    for _ in range(iterations):
        b = a.copy()
        b[1:] |= a[:-1]
        b[:-1] |= a[1:]
        b[:, 1:] |= a[:, :-1]
        b[:, :-1] |= a[:, 1:]
        a = b
    return a


def read_mask(mask: np.ndarray, length: int, flow_mask_dilates: int = MASK_DILATION, mask_dilates: int = MASK_DILATION):
    """P2 — read_mask for the ndarray the pipeline passes (propainter_inpaint.py:32-77): (flow_masks, masks_dilated), each a
    list of `length` uint8 arrays in {0, 255}; with 0 dilations the reference thresholds at 0.1 instead."""
    m = mask.squeeze(2) if mask.ndim == 3 and mask.shape[2] == 1 else mask
    if m.ndim != 2:
        raise ValueError("expected a single-channel mask")

    def one(it):
        return (binary_dilation_cross(m, it) if it > 0 else (m > 0.1)).astype(np.uint8) * 255

    return [one(flow_mask_dilates)] * length, [one(mask_dilates)] * length


def get_ref_index(mid: int, neighbor_ids: Sequence[int], length: int, ref_stride: int = REF_STRIDE, ref_num: int = -1) -> List[int]:
    """propainter_inpaint.py:120-135."""
    if ref_num == -1:
        return [i for i in range(0, length, ref_stride) if i not in neighbor_ids]
    out: List[int] = []
    for i in range(max(0, mid - ref_stride * (ref_num // 2)), min(length, mid + ref_stride * (ref_num // 2)), ref_stride):
        if i not in neighbor_ids:
            if len(out) > ref_num:
                break
            out.append(i)
    return out


def sub_ranges(length: int, sub: int, pad: int) -> List[Tuple[int, int, int, int]]:
    """The overlapped chunks a long sequence is processed in (flow completion :254-272 with pad 5, image propagation :284-304 with pad 10):
    chunk k covers [s, e) = [k*sub - pad, (k+1)*sub + pad) clipped to the sequence, and [keep_s, keep_e) of ITS result is what belongs to
    [k*sub, (k+1)*sub).  -> [(s, e, keep_s, keep_e)]"""
    out = []
    for f in range(0, length, sub):
        s, e = max(0, f - pad), min(length, f + sub + pad)
        out.append((s, e, f - s, (e - s) - (e - min(length, f + sub))))
    return out


def window_schedule(video_length: int, sub_video_length: int = 80) -> List[Tuple[List[int], List[int]]]:
    """P7 — the window loop (:318-333): (neighbor_ids, ref_ids) per window, a window every NEIGHBOR_LENGTH // 2 frames."""
    stride = NEIGHBOR_LENGTH // 2
    ref_num = sub_video_length // REF_STRIDE if video_length > sub_video_length else -1
    out = []
    for f in range(0, video_length, stride):
        nb = list(range(max(0, f - stride), min(video_length, f + stride + 1)))
        out.append((nb, get_ref_index(f, nb, video_length, REF_STRIDE, ref_num)))
    return out


def composite(comp_frames: List, pred_img: np.ndarray, binary_masks: np.ndarray, ori_frames: Sequence[np.ndarray], neighbor_ids: Sequence[int]):
    """P7 — :344-357: pred_img [n,h,w,3] float in 0..255, binary_masks [n,h,w,1] uint8 in {0,1}: truncate to u8, composite with the
    original, 0.5 / 0.5 blend with the previous visit, re-quantise to u8 after EVERY visit."""
    for i, idx in enumerate(neighbor_ids):
        img = np.array(pred_img[i]).astype(np.uint8) * binary_masks[i] + ori_frames[idx] * (1 - binary_masks[i])
        if comp_frames[idx] is None:
            comp_frames[idx] = img
        else:
            comp_frames[idx] = comp_frames[idx].astype(np.float32) * 0.5 + img.astype(np.float32) * 0.5
        comp_frames[idx] = comp_frames[idx].astype(np.uint8)
    return comp_frames
