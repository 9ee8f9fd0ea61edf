"""ProPainter on the B200: drop-in for backend/inpaint/propainter_inpaint.py `PropainterInpaint`.

STATUS — read this first: every stage is checked on the CPU against the oracle and the frames of the unmodified reference through
the fp32 stand-in of the device runtime (tests/test_propainter_cpu.py reproduces `PropainterInpaint.inpaint`'s golden output), and
all kernels compile for sm_100a, but the class has NOT run on a B200 yet (round 1's GPU budget was spent before it existed; DESIGN.md
§7).  It is therefore not exported from the package's top level, and tests/test_gpu_raft.py is gated on VSR_RUN_UNVALIDATED=1.

Stages (one shared device runtime): read_mask (host, propainter_tools) -> RaftFlow (P3, raft_flow.py) -> FlowCompletion (P4,
flow_completion.py) -> propagate_images (P5, flow_propagation.py) -> per window: Generator.encode_and_propagate + transform_and_decode
(P6, propainter_generator.py) -> composite / blend (P7, host, propainter_tools).  Clips longer than `sub_video_length` (the reference
chunks flow completion and propagation with overlaps there, propainter_inpaint.py:244-306) are not supported yet: `video_inpaint`
feeds batches of at most `propainterMaxLoadNum` frames, which is what `sub_video_length` is set to (main.py:171).
"""
import os
from typing import List, Sequence

import numpy as np

from . import _capi
from . import propainter_tools as PT
from .dbnet import _Tensor
from .flow_completion import FlowCompletion
from .flow_propagation import _Arena, propagate_images
from .propainter_generator import Generator, _GenRuntime
from .raft_flow import ITERS, RaftFlow


class PropainterInpaint:
    def __init__(self, device, model_dir, sub_video_length=80, use_fp16=True, runtime=None):
        """propainter_inpaint.py:139-190.  `use_fp16` is accepted for signature compatibility: the device path multiplies in fp16
        (fp32 accumulate) everywhere except the flow state, the correlation lookups and the FFT-free propagation arithmetic."""
        self.device, self.model_dir, self.sub_video_length = device, model_dir, sub_video_length
        self._rt = runtime if runtime is not None else _GenRuntime(device)
        self.fix_raft = RaftFlow(os.path.join(model_dir, "raft-things.pth"), runtime=self._rt)
        self.fix_flow_complete = FlowCompletion(os.path.join(model_dir, "recurrent_flow_completion.pth"), runtime=self._rt)
        self.model = Generator(os.path.join(model_dir, "ProPainter.pth"), runtime=self._rt)
        self._arena = _Arena(self._rt)
        self.raft_iter = ITERS

    def _flows(self, frames: Sequence[np.ndarray]):
        """propainter_inpaint.py:209-236: RAFT on clips of 12 / 8 / 4 / 2 frames by width, each clip overlapping the previous by one frame."""
        T, W = len(frames), frames[0].shape[1]
        clip = 12 if W <= 640 else 8 if W <= 720 else 4 if W <= 1280 else 2
        if T <= clip:
            return self.fix_raft(frames, self.raft_iter)
        ff, fb = [], []
        for f in range(0, T, clip):
            a, b = self.fix_raft(frames[max(f - 1, 0):min(T, f + clip)], self.raft_iter)
            ff.append(a)
            fb.append(b)
        return np.concatenate(ff), np.concatenate(fb)

    def inpaint(self, frames: Sequence[np.ndarray], mask: np.ndarray) -> List[np.ndarray]:
        """propainter_inpaint.py:192-361: BGR uint8 frames [H,W,3] (H, W multiples of 8, >= 128) + uint8 mask -> BGR uint8 frames."""
        frames = [np.ascontiguousarray(f, np.uint8) for f in frames]
        T, (H, W) = len(frames), frames[0].shape[:2]
        if T > self.sub_video_length:
            raise _capi.VsrError(f"{T} frames > sub_video_length {self.sub_video_length}: chunked completion / propagation is not supported yet")
        if T < 2:
            raise _capi.VsrError("ProPainter needs at least two frames (the reference routes single frames to LAMA, main.py:220)")
        rt = self._rt
        flow_masks, masks_dilated = PT.read_mask(mask, T)
        gf, gb = self._flows(frames)
        self._arena.begin(("inpaint", T, H, W))
        up = lambda arr: (lambda p: (rt.upload_to(p, arr), p)[1])(self._arena.alloc(max(np.ascontiguousarray(arr).nbytes, 16)))   # noqa: E731
        ff_dev, fb_dev = up(np.ascontiguousarray(gf, np.float32)), up(np.ascontiguousarray(gb, np.float32))
        fmask_dev, mask_dev = up(flow_masks[0]), up(masks_dilated[0])
        pf_dev, pb_dev = self.fix_flow_complete.complete(ff_dev, fb_dev, fmask_dev, T - 1, H, W)
        x = _Tensor(self._arena.alloc(T * H * W * 8 * 2), 3, H, W, 8, n=T)
        rt.frames(frames, x)
        state = propagate_images(rt, x, mask_dev, pf_dev, pb_dev, self._arena)
        comp: List = [None] * T
        binary = (masks_dilated[0] > 0).astype(np.uint8)[None, :, :, None]
        rgb = [np.ascontiguousarray(f[:, :, ::-1]) for f in frames]
        for nb, refs in PT.window_schedule(T, self.sub_video_length):
            ids = nb + refs
            enc, _ = self.model.encode_and_propagate(state, mask_dev, ids, pf_dev, pb_dev, len(nb))
            pred = self.model.transform_and_decode(enc, len(nb), masks_dilated[0], H, W)
            PT.composite(comp, pred, np.repeat(binary, len(nb), 0), rgb, nb)
        return [np.ascontiguousarray(c[:, :, ::-1]) for c in comp]

    def __call__(self, input_frames: List[np.ndarray], input_mask: np.ndarray) -> List[np.ndarray]:
        """propainter_inpaint.py:363-418: strips (heights multiples of 8) at native resolution, the first frame's mask for the whole batch,
        every strip replaced whole; the input frames are not modified."""
        mask = input_mask if input_mask.ndim == 2 else input_mask[:, :, 0]
        H, W = mask.shape[:2]
        out = [f.copy() for f in input_frames]
        for (y0, y1, x0, x1) in PT.strip_areas(W, H, mask):
            comps = self.inpaint([f[y0:y1, x0:x1] for f in out], mask[y0:y1, x0:x1])
            for f, c in zip(out, comps):
                f[y0:y1, x0:x1] = c
        return out


__all__ = ["PropainterInpaint"]
