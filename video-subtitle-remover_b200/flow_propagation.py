"""ProPainter's flow-guided image propagation on the device runtime.

STATUS: like raft_flow.py — checked against the oracle (oracle/propainter_oracle.img_propagation, pinned to the reference's taps)
through the CPU stand-in of the runtime (tests/test_propagation_cpu.py); its kernels compile for sm_100a; NOT yet run on a B200.

Mirrors `InpaintGenerator.img_propagation` = `BidirectionalPropagation(3, learnable=False)` (video/model/propainter.py:107-193,316-319)
and the composition around it (propainter_inpaint.py:283,308-312).  A frame and its mask travel together as an 8-half pixel
(channels 0..2 frame in [-1,1], channel 3 mask), so one propagation step is ONE launch per frame: consistency check of the two
flows, nearest warp of the previous frame, bilinear warp of the previous mask and the masked update, all per pixel.
"""
import ctypes as C
from typing import Tuple

import numpy as np

from . import _capi
from .dbnet import _Tensor
from .raft_flow import _RaftRuntime


class _PropRuntime(_RaftRuntime):
    def img_prop_step(self, prev, cur, flow_prop, flow_check, out):
        _capi.check(self.L.vsr_rt_img_prop_step(self.h, prev.ptr, cur.ptr, flow_prop, flow_check, cur.h, cur.w, out.ptr))

    def prop_state(self, frames, mask_u8, prop, out):
        _capi.check(self.L.vsr_rt_prop_state(self.h, frames.ptr, mask_u8, prop.ptr if prop is not None else 0, frames.n, frames.h, frames.w, out.ptr))

    def copy_bytes(self, src: int, dst: int, nbytes: int):
        """device -> device, in stream order"""
        _capi.check(self.L.vsr_rt_copy(self.h, dst, src, nbytes))

    def upload_bytes(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr)
        p = self.alloc(max(arr.nbytes, 16))
        self.upload_to(p, arr)
        return p

    def upload_to(self, ptr: int, arr: np.ndarray):
        """host array -> an existing device buffer (buffers live as long as the runtime: per-call inputs go through an _Arena)"""
        arr = np.ascontiguousarray(arr)
        _capi.check(self.L.vsr_rt_upload(self.h, ptr, arr.ctypes.data_as(C.c_void_p), arr.nbytes))


class _Arena:
    """Replays the allocation sequence of an eager graph: the first call with a key allocates, later calls with the same key get the
    same buffers back in the same order (the runtime frees nothing before it is destroyed, and zero-initialises only fresh buffers —
    every launch sequence here rewrites exactly the regions it wrote the first time)."""

    KEEP = 1   # epochs a shape key survives without being used

    def __init__(self, rt):
        self.rt, self.seqs, self.cur, self.i = rt, {}, None, 0
        self.epoch, self.used = 0, {}

    def tick(self):
        """A new top-level call (one batch of `propainter_mode`, main.py:229-241) starts.  Batch lengths differ per interval and every
        length costs GBs of work buffers at a 1080p strip, while the runtime itself frees nothing before it is destroyed: release the
        buffers of the shape keys that the last KEEP + 1 calls did not touch.  Nothing of an earlier call is live at this point."""
        self.epoch += 1
        for key in [k for k, e in self.used.items() if e < self.epoch - self.KEEP]:
            for _, ptr in self.seqs.pop(key):
                self.rt.free(ptr)
            del self.used[key]

    def begin(self, key):
        self.cur, self.i = self.seqs.setdefault(key, []), 0
        self.used[key] = self.epoch

    def alloc(self, nbytes: int) -> int:
        if self.i < len(self.cur):
            size, ptr = self.cur[self.i]
            if size != nbytes:
                raise _capi.VsrError("allocation sequence changed between calls with the same shape key")
        else:
            ptr = self.rt.alloc(nbytes)
            self.cur.append((nbytes, ptr))
        self.i += 1
        return ptr


def _image(t: _Tensor, k: int) -> _Tensor:
    """image k of a [n,h,w,cp] tensor as a 1-image tensor"""
    return _Tensor(t.ptr + k * t.h * t.w * t.cp * 2, t.c, t.h, t.w, t.cp, n=1)


def propagate_images(rt, frames: _Tensor, mask_u8: int, flows_f: int, flows_b: int, arena: "_Arena" = None) -> _Tensor:
    """frames: fp16 [T,H,W,8] RGB in [-1,1] (vsr_rt_pp_frames); mask_u8: device u8 [H,W] (the dilated mask, > 0 = hole);
    flows_f / flows_b: device fp32 [T-1,2,H,W] (completed flows t -> t+1 and t+1 -> t).
    Returns the state tensor [T,H,W,8]: channels 0..2 = frame*(1-m) + propagated*m (`updated_frames`), channel 3 = `updated_masks`."""
    T, H, W = frames.n, frames.h, frames.w
    alloc = rt.alloc
    if arena is not None:
        arena.begin(("prop", T, H, W))
        alloc = arena.alloc
    new = lambda: _Tensor(alloc(T * H * W * 8 * 2), 4, H, W, 8, n=T)   # noqa: E731
    s0, sb, sf, out = new(), new(), new(), new()
    fl = lambda base, i: base + i * 2 * H * W * 4                         # noqa: E731
    rt.prop_state(frames, mask_u8, None, s0)
    rt.copy_channels(_image(s0, T - 1), _image(sb, T - 1), 0, 8)
    for idx in range(T - 2, -1, -1):           # backward pass: frame idx takes from idx + 1 along the forward flow
        rt.img_prop_step(_image(sb, idx + 1), _image(s0, idx), fl(flows_f, idx), fl(flows_b, idx), _image(sb, idx))
    rt.copy_channels(_image(sb, 0), _image(sf, 0), 0, 8)
    for idx in range(1, T):                    # forward pass over the backward pass's result
        rt.img_prop_step(_image(sf, idx - 1), _image(sb, idx), fl(flows_b, idx - 1), fl(flows_f, idx - 1), _image(sf, idx))
    rt.prop_state(frames, mask_u8, sf, out)
    return out


def propagate_images_host(rt, frames_bgr, mask: np.ndarray, flows_f: np.ndarray, flows_b: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Host-array convenience around `propagate_images` (tests, stage-level use): BGR uint8 frames, uint8 mask [H,W], float32 flows
    [T-1,2,H,W] -> (updated frames float32 [T,3,H,W] in [-1,1], updated masks float32 [T,1,H,W])."""
    frames_bgr = [np.ascontiguousarray(f, np.uint8) for f in frames_bgr]
    T, (H, W) = len(frames_bgr), frames_bgr[0].shape[:2]
    x = _Tensor(rt.alloc(T * H * W * 8 * 2), 3, H, W, 8, n=T)
    rt.frames(frames_bgr, x)
    out = propagate_images(rt, x, rt.upload_bytes(np.ascontiguousarray(mask, np.uint8)), rt.upload_bytes(np.ascontiguousarray(flows_f, np.float32)),
                           rt.upload_bytes(np.ascontiguousarray(flows_b, np.float32)))
    host = rt.download(_Tensor(out.ptr, 8, T * H, W, 8)).astype(np.float32).reshape(T, H, W, 8)
    return np.ascontiguousarray(host[..., :3].transpose(0, 3, 1, 2)), np.ascontiguousarray(host[..., 3:4].transpose(0, 3, 1, 2))


__all__ = ["propagate_images", "propagate_images_host"]
