"""Host-side mirror of backend/inpaint/sttn_det_inpaint.py on the same C-ABI
engine as STTNInpaint, created in mode 1: 432x240 model input, 108x60 feature map, patches
(108,60) (36,20) (18,10) (9,5); the strip of the (un-thresholded) mask is resized like the frames, gates the
encoder input (mask/255 > 0.5) and the low-res composite (mask > 0), and the whole strip of the output frame
is replaced by the up-scaled comp (sttn_det_inpaint.py:93)."""
import ctypes as C
from typing import List

import numpy as np

from . import _capi
from .sttn_auto_inpaint import STTNInpaint


class STTNDetInpaint(STTNInpaint):
    """Drop-in for backend/inpaint/sttn_det_inpaint.py:23 `STTNDetInpaint(device, model_path)`.

    `model(frames, mask)` keeps the reference convention: BGR uint8 frames [H,W,3], mask [H,W] uint8 (0/255,
    as built by create_mask) -> new frames; `video_inpaint` (backend/main.py:326) calls it per batch."""

    _DET = True

    def inpaint(self, frames: List[np.ndarray], masks: List[np.ndarray]):
        """sttn_det_inpaint.py:124-174 on already-scaled frames [240,432,3] BGR and masks [240,432] (0..255).
        The reference receives one resized mask per frame, all identical (:66-75); the first one is used."""
        T = len(frames)
        x = np.ascontiguousarray(np.stack(frames), dtype=np.uint8)
        m = np.ascontiguousarray(np.asarray(masks[0]).reshape(self.model_input_height, self.model_input_width), dtype=np.uint8)
        if x.shape[1:] != (self.model_input_height, self.model_input_width, 3):
            raise ValueError(f"strip frames must be {(self.model_input_height, self.model_input_width, 3)}, got {x.shape[1:]}")
        comps = np.empty(x.shape, np.float32)
        visits = np.zeros(T, np.int32)
        self._exact_softmax_retry(lambda: _capi.check(_capi.lib().vsr_sttn_inpaint_strip_masked(
            self._h, _capi.ptr(x, C.c_uint8), _capi.ptr(m, C.c_uint8), T, _capi.ptr(comps, C.c_float), _capi.ptr(visits, C.c_int32))))
        return [comps[i].astype(np.uint8) if visits[i] <= 1 else comps[i] for i in range(T)]
