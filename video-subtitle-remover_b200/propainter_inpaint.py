"""ProPainter on the B200: drop-in for backend/inpaint/propainter_inpaint.py `PropainterInpaint`.

Validated on a B200 in round 2 (tests/test_gpu_raft.py: RAFT end-point error, image propagation, flow completion and the whole pipeline
against the frames of the unmodified reference, >= 45 dB in the hole, bit-exact outside; profiles/gpu_session_r2_s1_summary.txt).

Stages (one shared device runtime): read_mask (host, propainter_tools) -> RaftFlow (P3, raft_flow.py) -> FlowCompletion (P4,
flow_completion.py) -> propagate_images (P5, flow_propagation.py) -> per window: Generator.encode_and_propagate + transform_and_decode
(P6, propainter_generator.py) -> composite / blend (P7, host, propainter_tools).  Clips longer than `sub_video_length` go through the
reference's overlapped chunks of flow completion and image propagation and its capped reference frames (propainter_inpaint.py:251-324);
`video_inpaint` itself feeds batches of at most `propainterMaxLoadNum` frames, which is what `sub_video_length` is set to (main.py:171).
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


def flow_clips(n_frames: int, width: int, clip=None) -> List[tuple]:
    """(first, end) frame ranges RAFT runs on (propainter_inpaint.py:209-236): clips of 12 / 8 / 4 / 2 frames by width, every clip after the
    first starting one frame early so that consecutive clips share the pair in between."""
    clip = clip or (12 if width <= 640 else 8 if width <= 720 else 4 if width <= 1280 else 2)
    if n_frames <= clip:
        return [(0, n_frames)]
    return [(max(f - 1, 0), min(n_frames, f + clip)) for f in range(0, n_frames, clip)]


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
        self.raft_clip = None          # frames per RAFT clip; None = by width like the reference (:209-216)

    def _flows(self, frames: Sequence[np.ndarray], shard=None):
        """propainter_inpaint.py:209-236: RAFT on clips of 12 / 8 / 4 / 2 frames by width, each clip overlapping the previous by one frame.
        Clips are independent (a flow depends on its two frames only): with a `shard` every rank runs RAFT on its share of the clips and
        the flows are exchanged, so that flow completion sees the whole sequence on every rank."""
        clips = flow_clips(len(frames), frames[0].shape[1], self.raft_clip)
        mine = {i: np.stack(self.fix_raft(frames[a:b], self.raft_iter)) for i, (a, b) in enumerate(clips) if shard is None or shard.owns(i)}
        parts = mine if shard is None else shard.exchange(mine)
        return np.concatenate([parts[i][0] for i in range(len(clips))]), np.concatenate([parts[i][1] for i in range(len(clips))])

    def inpaint(self, frames: Sequence[np.ndarray], mask: np.ndarray, shard=None) -> List[np.ndarray]:
        """propainter_inpaint.py:192-361: BGR uint8 frames [H,W,3] (H, W multiples of 8, >= 128) + uint8 mask -> BGR uint8 frames.

        `shard` (vsr_b200.distributed.Shard) splits one sub-video over the ranks of a process group (BASELINE config 5): RAFT clips and
        generator windows are dealt round-robin; flow completion and image propagation are sequential over the frames and run replicated;
        two exchanges (flows after RAFT, the u8 window predictions before the blend) make every rank return the same frames, identical
        to the unsharded result — windows overlap and P7's 0.5 / 0.5 blend is order dependent, so the blend always runs in schedule order."""
        frames = [np.ascontiguousarray(f, np.uint8) for f in frames]
        T, (H, W) = len(frames), frames[0].shape[:2]
        if T < 2:
            raise _capi.VsrError("ProPainter needs at least two frames (the reference routes single frames to LAMA, main.py:220)")
        rt = self._rt
        for arena in (self._arena, self.fix_flow_complete._arena, self.model._arena):
            arena.tick()
        import time

        clock = {"t": time.perf_counter()}
        self.stage_seconds = getattr(self, "stage_seconds", None) or {}

        def lap(name):      # wall clock per stage (the stages end in a host read or an explicit sync): bench.py reports the last call's
            rt.sync()
            now = time.perf_counter()
            self.stage_seconds[name] = self.stage_seconds.get(name, 0.0) + now - clock["t"]
            clock["t"] = now

        self.stage_seconds.clear()
        flow_masks, masks_dilated = PT.read_mask(mask, T)
        lap("read_mask")
        N, fbytes, sbytes = T - 1, 2 * H * W * 4, H * W * 8 * 2         # flows, bytes of one flow field / of one frame of the state tensor
        self._arena.begin(("inpaint", T, H, W))
        up = lambda arr: (lambda p: (rt.upload_to(p, arr), p)[1])(self._arena.alloc(max(np.ascontiguousarray(arr).nbytes, 16)))   # noqa: E731
        if shard is None:                                                   # RAFT's flows go straight into the completion network's input
            ff_dev, fb_dev = self._arena.alloc(N * fbytes), self._arena.alloc(N * fbytes)
            for a, b in flow_clips(T, W, self.raft_clip):                   # clip (a, b) holds the pairs a .. b-2 (clips overlap by one frame)
                self.fix_raft(frames[a:b], self.raft_iter, dst=(ff_dev + a * fbytes, fb_dev + a * fbytes))
        else:                                                               # sharded: the clips' flows are exchanged between the ranks
            gf, gb = self._flows(frames, shard)
            ff_dev, fb_dev = up(np.ascontiguousarray(gf, np.float32)), up(np.ascontiguousarray(gb, np.float32))
        lap("raft")
        fmask_dev, mask_dev = up(flow_masks[0]), up(masks_dilated[0])
        if N > self.sub_video_length:                                       # :251-276: overlapped chunks (pad 5), the middle of each kept
            pf_dev, pb_dev = self._arena.alloc(N * fbytes), self._arena.alloc(N * fbytes)
            for s, e, ks, ke in PT.sub_ranges(N, self.sub_video_length, 5):
                a, b = self.fix_flow_complete.complete(ff_dev + s * fbytes, fb_dev + s * fbytes, fmask_dev, e - s, H, W)
                rt.copy_bytes(a + ks * fbytes, pf_dev + (s + ks) * fbytes, (ke - ks) * fbytes)   # before the next chunk reuses the buffers
                rt.copy_bytes(b + ks * fbytes, pb_dev + (s + ks) * fbytes, (ke - ks) * fbytes)
        else:
            pf_dev, pb_dev = self.fix_flow_complete.complete(ff_dev, fb_dev, fmask_dev, N, H, W)
        lap("flow_completion")
        x = _Tensor(self._arena.alloc(T * sbytes), 3, H, W, 8, n=T)
        rt.frames(frames, x)
        sub_prop = min(100, self.sub_video_length)
        if T > sub_prop:                                                    # :281-312: chunks of min(100, sub_video_length) frames, pad 10
            state = _Tensor(self._arena.alloc(T * sbytes), 4, H, W, 8, n=T)
            for s, e, ks, ke in PT.sub_ranges(T, sub_prop, 10):
                part = propagate_images(rt, _Tensor(x.ptr + s * sbytes, 3, H, W, 8, n=e - s), mask_dev, pf_dev + s * fbytes, pb_dev + s * fbytes, self._arena)
                rt.copy_bytes(part.ptr + ks * sbytes, state.ptr + (s + ks) * sbytes, (ke - ks) * sbytes)
        else:
            state = propagate_images(rt, x, mask_dev, pf_dev, pb_dev, self._arena)
        lap("image_propagation")
        comp: List = [None] * T
        binary = (masks_dilated[0] > 0).astype(np.uint8)[None, :, :, None]
        rgb = [np.ascontiguousarray(f[:, :, ::-1]) for f in frames]
        schedule = PT.window_schedule(T, self.sub_video_length)
        preds = {}
        for wi, (nb, refs) in enumerate(schedule):
            if shard is not None and not shard.owns(wi):
                continue
            enc, _ = self.model.encode_and_propagate(state, mask_dev, nb + refs, pf_dev, pb_dev, len(nb))
            pred = self.model.transform_and_decode(enc, len(nb), masks_dilated[0], H, W)
            if shard is None:
                PT.composite(comp, pred, np.repeat(binary, len(nb), 0), rgb, nb)
            else:
                preds[wi] = pred
        if shard is not None:
            preds = shard.exchange(preds)
            for wi, (nb, _) in enumerate(schedule):
                PT.composite(comp, preds[wi], np.repeat(binary, len(nb), 0), rgb, nb)
        lap("generator_and_composite")
        return [np.ascontiguousarray(c[:, :, ::-1]) for c in comp]

    def __call__(self, input_frames: List[np.ndarray], input_mask: np.ndarray, shard=None) -> List[np.ndarray]:
        """propainter_inpaint.py:363-418: strips (heights multiples of 8) at native resolution, the first frame's mask for the whole batch,
        every strip replaced whole; the input frames are not modified."""
        mask = input_mask if input_mask.ndim == 2 else input_mask[:, :, 0]
        H, W = mask.shape[:2]
        out = [f.copy() for f in input_frames]
        for (y0, y1, x0, x1) in PT.strip_areas(W, H, mask):
            comps = self.inpaint([f[y0:y1, x0:x1] for f in out], mask[y0:y1, x0:x1], shard)
            for f, c in zip(out, comps):
                f[y0:y1, x0:x1] = c
        return out


__all__ = ["PropainterInpaint", "flow_clips"]
