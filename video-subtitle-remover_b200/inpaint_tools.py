"""Integer mask / index path with the reference's function names (backend/tools/inpaint_tools.py),
computed by the host C++ in csrc/host_index.h through the C ABI (bit-exact, no GPU needed)."""
import ctypes as C

import numpy as np

from . import _capi
from .config import config


def create_mask(size, coords_list):
    """backend/tools/inpaint_tools.py:31-47.  size=(H,W); coords=(xmin,xmax,ymin,ymax) -> uint8 {0,255}."""
    H, W = int(size[0]), int(size[1])
    mask = np.zeros((H, W), dtype=np.uint8)
    boxes = np.asarray(coords_list if coords_list else [], dtype=np.int32).reshape(-1, 4)
    _capi.check(_capi.lib().vsr_create_mask(_capi.ptr(mask, C.c_uint8), H, W, _capi.ptr(boxes, C.c_int32), len(boxes),
                                            int(config.subtitleAreaDeviationPixel.value)))
    return mask


def get_inpaint_area_by_mask(W, H, h, mask, multiple=1):
    """backend/tools/inpaint_tools.py:49-242 -> [(ymin, ymax, xmin, xmax), ...]."""
    m = np.asarray(mask)
    if m.ndim == 3:
        m = m[:, :, 0]
    m = np.ascontiguousarray((m > 0).astype(np.uint8))
    if m.shape != (H, W):
        raise ValueError(f"mask shape {m.shape} != {(H, W)}")
    cap = max(16, H)
    while True:     # the reference returns however many areas the mask produces: grow the buffer until they fit
        out = np.zeros((cap, 4), dtype=np.int32)
        try:
            n = _capi.check(_capi.lib().vsr_inpaint_area_by_mask(int(W), int(H), int(h), _capi.ptr(m, C.c_uint8), int(multiple),
                                                                 _capi.ptr(out, C.c_int32), cap))
        except _capi.VsrError as e:
            if "areas buffer too small" not in str(e) or cap > 64 * max(16, H * W):
                raise
            cap *= 4
            continue
        return [tuple(int(v) for v in out[i]) for i in range(n)]


def batch_generator(data, max_batch_size):
    """backend/tools/inpaint_tools.py:7-29 (generator of slices of `data`)."""
    n = len(data)
    sizes = np.zeros(max(n, 1) + 1, dtype=np.int32)
    k = _capi.check(_capi.lib().vsr_batch_sizes(n, int(max_batch_size), _capi.ptr(sizes, C.c_int32), len(sizes)))
    start = 0
    for i in range(k):
        yield data[start:start + int(sizes[i])]
        start += int(sizes[i])


def window_schedule(T, stride, ref_length):
    """sttn_auto_inpaint.py:142-146 + get_ref_index :107-120 -> [(neighbor_ids, ref_ids)]."""
    maxw = T // max(stride, 1) + 2
    per = T + 1
    ids = np.zeros((maxw, per), dtype=np.int32)
    nn = np.zeros(maxw, dtype=np.int32)
    nr = np.zeros(maxw, dtype=np.int32)
    k = _capi.check(_capi.lib().vsr_window_schedule(int(T), int(stride), int(ref_length), _capi.ptr(ids, C.c_int32),
                                                    _capi.ptr(nn, C.c_int32), _capi.ptr(nr, C.c_int32), maxw, per))
    return [(ids[w, :nn[w]].tolist(), ids[w, nn[w]:nn[w] + nr[w]].tolist()) for w in range(k)]
