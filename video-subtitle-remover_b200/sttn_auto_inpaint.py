"""Host-side mirror of backend/inpaint/sttn_auto_inpaint.py on top of the C-ABI engine.

`STTNInpaint` and `STTNAutoInpaint` keep the reference's constructor signatures, call conventions and
hyper-parameter sources (config.sttnNeighborStride / sttnReferenceLength / getSttnMaxLoadNum read at
construction, sttn_auto_inpaint.py:40-41,195); everything between "list of BGR frames + mask" and
"list of BGR frames" runs on the B200 through include/vsr_b200.h.  There is no CPU path.
"""
import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Union

import numpy as np

from . import _capi
from .config import config
from .inpaint_tools import get_inpaint_area_by_mask, window_schedule


def _device_index(device) -> int:
    """Accepts torch.device('cuda:0') (what HardwareAccelerator.device hands out), 'cuda:1', or an int."""
    if isinstance(device, int):
        return device
    idx = getattr(device, "index", None)
    typ = getattr(device, "type", None)
    if typ is not None:
        if typ != "cuda":
            raise _capi.VsrError(f"vsr_b200 runs on CUDA sm_100a only; got device '{device}' (no CPU fallback)")
        return 0 if idx is None else int(idx)
    s = str(device)
    if s.startswith("cuda"):
        return int(s.split(":")[1]) if ":" in s else 0
    raise _capi.VsrError(f"vsr_b200 runs on CUDA sm_100a only; got device '{device}' (no CPU fallback)")


def load_state_dict(model_path) -> Dict[str, np.ndarray]:
    """ckpt['netG'] of the reference checkpoint (sttn_auto_inpaint.py:34), an .npz, or a ready dict."""
    if isinstance(model_path, dict):
        sd = model_path
    elif str(model_path).endswith(".npz"):
        z = np.load(model_path)
        sd = {k: z[k] for k in z.files}
    else:
        import torch  # unpickling only

        sd = torch.load(model_path, map_location="cpu", weights_only=False)
        sd = sd["netG"] if "netG" in sd else sd
    out = {}
    for k, v in sd.items():
        if hasattr(v, "detach"):
            v = v.detach().cpu().float().numpy()
        out[k] = np.ascontiguousarray(v, dtype=np.float32)
    return out


class STTNInpaint:
    """Drop-in for backend/inpaint/sttn_auto_inpaint.py:28 `STTNInpaint`."""

    _DET = False  # STTNDetInpaint flips this: 432x240 geometry, masked encoder input, low-res composite

    def __init__(self, device, model_path):
        self.device = device
        self._dev = _device_index(device)
        self.neighbor_stride = config.sttnNeighborStride.value      # :40
        self.ref_length = config.sttnReferenceLength.value          # :41
        L = _capi.lib()
        cfg = _capi.Config()
        (L.vsr_sttn_det_config if self._DET else L.vsr_sttn_default_config)(C.byref(cfg))
        self.model_input_width, self.model_input_height = int(cfg.model_w), int(cfg.model_h)  # :38 / sttn_det_inpaint.py:33
        cfg.neighbor_stride = int(self.neighbor_stride)
        cfg.ref_length = int(self.ref_length)
        h = C.c_void_p()
        _capi.check(L.vsr_sttn_create(C.byref(h), self._dev, C.byref(cfg)))
        self._h = h
        for name, arr in load_state_dict(model_path).items():
            if not (name.startswith("encoder.") or name.startswith("decoder.") or name.startswith("transformer.")):
                continue
            shape = np.asarray(arr.shape, dtype=np.int64)
            _capi.check(L.vsr_sttn_set_weight(self._h, name.encode(), _capi.ptr(arr, C.c_float), _capi.ptr(shape, C.c_int64),
                                              arr.ndim))
        _capi.check(L.vsr_sttn_finalize_weights(self._h))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _capi.lib().vsr_sttn_destroy(h)
            except Exception:
                pass
            self._h = None

    # ---- reference API -------------------------------------------------------------------------
    def __call__(self, input_frames: List[np.ndarray], input_mask: np.ndarray) -> List[np.ndarray]:
        """sttn_auto_inpaint.py:43-97: BGR uint8 frames [H,W,3] + mask [H,W] (0/255) -> new frames."""
        if len(input_frames) == 0:
            return []
        outs = [np.empty_like(np.ascontiguousarray(f, dtype=np.uint8)) for f in input_frames]
        self._run(input_frames, input_mask, outs)
        return outs

    def inpaint_inplace(self, frames: Sequence[np.ndarray], input_mask: np.ndarray) -> None:
        """Same computation writing the result strips into `frames` themselves (what the chunk loop of
        STTNAutoInpaint.__call__ does with `frames_hr`, sttn_auto_inpaint.py:299-315)."""
        if len(frames):
            self._run(frames, input_mask, frames)

    def inpaint(self, frames: List[np.ndarray]):
        """sttn_auto_inpaint.py:122-164 on already-scaled strip frames [120,640,3] BGR.  Returns comps in
        RGB: uint8 for frames decoded once, float32 (running 0.5/0.5 blend) otherwise."""
        T = len(frames)
        x = np.ascontiguousarray(np.stack(frames), dtype=np.uint8)
        if x.shape[1:] != (self.model_input_height, self.model_input_width, 3):
            raise ValueError(f"strip frames must be {(self.model_input_height, self.model_input_width, 3)}, got {x.shape[1:]}")
        comps = np.empty(x.shape, np.float32)
        visits = np.zeros(T, np.int32)
        self._exact_softmax_retry(lambda: _capi.check(_capi.lib().vsr_sttn_inpaint_strip(self._h, _capi.ptr(x, C.c_uint8), T, _capi.ptr(comps, C.c_float),
                                                                                          _capi.ptr(visits, C.c_int32))))
        return [comps[i].astype(np.uint8) if visits[i] <= 1 else comps[i] for i in range(T)]

    def get_ref_index(self, neighbor_ids, length):
        """sttn_auto_inpaint.py:107-120."""
        return [i for i in range(0, length, self.ref_length) if i not in neighbor_ids]

    @staticmethod
    def read_mask(path):
        """sttn_auto_inpaint.py:99-105."""
        import cv2

        img = cv2.imread(path, 0)
        _, img = cv2.threshold(img, 127, 1, cv2.THRESH_BINARY)
        return img[:, :, None]

    # ---- engine plumbing -----------------------------------------------------------------------
    def _ptr_array(self, frames):
        arr = (C.c_void_p * len(frames))()
        keep = []
        for i, f in enumerate(frames):
            if not (isinstance(f, np.ndarray) and f.dtype == np.uint8 and f.flags["C_CONTIGUOUS"]):
                raise ValueError("frames must be C-contiguous uint8 arrays")
            keep.append(f)
            arr[i] = f.ctypes.data
        return arr, keep

    def _check_batch(self, frames, mask):
        H, W = frames[0].shape[:2]
        for f in frames:
            if f.shape != (H, W, 3):
                raise ValueError("all frames must share one [H,W,3] shape")
        m = np.asarray(mask)
        if m.ndim == 3:
            m = m[:, :, 0]
        if m.shape != (H, W):
            raise ValueError(f"mask shape {m.shape} != frame shape {(H, W)}")
        return H, W, np.ascontiguousarray(m, dtype=np.uint8)

    def set_option(self, name: str, value: int) -> None:
        _capi.check(_capi.lib().vsr_sttn_set_option(self._h, name.encode(), int(value)))

    def _exact_softmax_retry(self, call):
        """Run `call()`; when the single-pass softmax reports logits beyond its range (never seen with the reference's weights: it would take
        |logit| > 69), switch this engine to the score-matrix + softmax-kernel path for good and repeat the call."""
        try:
            return call()
        except _capi.VsrRangeError:
            self.set_option("attn_direct", 0)
            return call()

    def _run(self, frames_in, mask, frames_out):
        return self._exact_softmax_retry(lambda: self._run_once(frames_in, mask, frames_out))

    def _run_once(self, frames_in, mask, frames_out):
        frames_in = [np.ascontiguousarray(f, dtype=np.uint8) for f in frames_in]
        H, W, m = self._check_batch(frames_in, mask)
        pin, keep_in = self._ptr_array(frames_in)
        pout, keep_out = self._ptr_array(list(frames_out))
        L = _capi.lib()
        _capi.check(L.vsr_sttn_inpaint_frames(self._h, C.cast(pin, C.POINTER(C.c_void_p)), len(frames_in), H, W,
                                              _capi.ptr(m, C.c_uint8), C.cast(pout, C.POINTER(C.c_void_p))))

    # split form used by bench.py
    def stage(self, frames_in, mask):
        frames_in = [np.ascontiguousarray(f, dtype=np.uint8) for f in frames_in]
        H, W, m = self._check_batch(frames_in, mask)
        pin, keep = self._ptr_array(frames_in)
        self._staged = (pin, keep, m)
        _capi.check(_capi.lib().vsr_sttn_stage(self._h, C.cast(pin, C.POINTER(C.c_void_p)), len(frames_in), H, W,
                                               _capi.ptr(m, C.c_uint8)))

    def compute(self):
        _capi.check(_capi.lib().vsr_sttn_compute(self._h))

    def fetch(self, frames_out):
        pout, keep = self._ptr_array(list(frames_out))
        _capi.check(_capi.lib().vsr_sttn_fetch(self._h, C.cast(pout, C.POINTER(C.c_void_p))))

    # asynchronous chunk pipeline (two chunks in flight)
    def submit(self, frames, mask) -> int:
        """Enqueue one chunk (strips H2D, kernels, D2H) and return a ticket; `frames` must stay alive and
        unmodified until `collect(ticket, ...)`."""
        frames = list(frames)
        H, W, m = self._check_batch(frames, mask)
        pin, keep = self._ptr_array(frames)
        t = int(_capi.lib().vsr_sttn_submit(self._h, C.cast(pin, C.POINTER(C.c_void_p)), len(frames), H, W,
                                            _capi.ptr(m, C.c_uint8)))
        _capi.check(t)
        self._inflight = getattr(self, "_inflight", {})
        self._inflight[t] = (keep, m)
        return t

    def collect(self, ticket: int, frames_out) -> None:
        frames_out = list(frames_out)
        pout, keep = self._ptr_array(frames_out)
        suspect = getattr(self, "_suspect", set())
        src, m = self._inflight[ticket]
        in_place = any(a is b for a, b in zip(src, frames_out))
        try:
            if ticket in suspect and in_place:     # its result must not land in the frames the repeat reads: collect into scratch copies
                scratch, _ = self._ptr_array([f.copy() for f in src])
                _capi.check(_capi.lib().vsr_sttn_collect(self._h, ticket, C.cast(scratch, C.POINTER(C.c_void_p))))
            else:
                _capi.check(_capi.lib().vsr_sttn_collect(self._h, ticket, C.cast(pout, C.POINTER(C.c_void_p))))
            redo = ticket in suspect
        except _capi.VsrRangeError:      # repeat this chunk synchronously on the exact-softmax path (its input frames are still untouched);
            self.set_option("attn_direct", 0)   # a chunk submitted before the switch shares the (now cleared) flag: repeat it as well
            self._suspect = suspect = suspect | (set(self._inflight) - {ticket})
            redo = True
        if redo:
            self._run_once(src, m, frames_out)
        suspect.discard(ticket)
        self._inflight.pop(ticket, None)

    def sync(self):
        _capi.check(_capi.lib().vsr_sttn_sync(self._h))

    # ---- one chunk over several GPUs: window-level sharding (single-clip strong scaling) --------------------
    def inpaint_chunk_sharded(self, frames: Sequence[np.ndarray], input_mask: np.ndarray, rank: int, world: int, all_gather=None) -> List[int]:
        """The windows of one chunk's schedule (sttn_auto_inpaint.py:142-146) dealt over the `world` ranks of a process group: window w runs
        on rank w % world.  Two all-gathers over NVLink carry the only data that crosses windows — the encoder features of the
        chunk's reference frames (get_ref_index, :107-120) and, afterwards, the windows' quantised predictions for the ordered 0.5 / 0.5
        blend (:159-162) — both on device memory (`all_gather(device_pointer, region_bytes)`, default: NCCL through torch.distributed
        on device tensors filled by device-to-device copies).  Every rank must call this with the same frames and mask.  The result strips of the frames f with
        f % world == rank are written into `frames[f]` in place (the other frames are left as they are on this rank); returns those
        frame indices.  Window by window the arithmetic is the single-GPU one and the blend is replayed in schedule order; the output
        differs from the unsharded call only through the summation order of the split-K attention heads, which depends on which windows
        share a launch (measured: one grey level on about 1 % of the pixels, never more than two; with world = 1 bit-identical)."""
        frames = list(frames)
        if not frames:
            return []
        ref_ptr, ref_bytes, pred_ptr, pred_bytes = self.shard_begin(frames, input_mask, rank, world)
        gather = all_gather or (lambda ptr, nbytes: _nccl_all_gather(self, ptr, nbytes, rank, world))
        if world > 1:
            gather(ref_ptr, ref_bytes)
        self.shard_windows()
        if world > 1:
            gather(pred_ptr, pred_bytes)
        try:
            self.shard_finish(frames)
        except _capi.VsrRangeError:
            if getattr(self, "_in_retry", False):
                raise
            self.set_option("attn_direct", 0)      # the flags travelled with the predictions: every rank takes this branch together
            self._in_retry = True
            try:
                return self.inpaint_chunk_sharded(frames, input_mask, rank, world, all_gather)
            finally:
                self._in_retry = False
        return list(range(rank, len(frames), world))

    # the three phases of the sharded chunk (include/vsr_b200.h: vsr_sttn_shard_*); between them the caller all-gathers the two buffers
    def shard_begin(self, frames: Sequence[np.ndarray], input_mask: np.ndarray, rank: int, world: int):
        """-> (reference-feature buffer pointer, bytes per rank region, prediction buffer pointer, bytes per rank region)"""
        H, W, m = self._check_batch(frames, input_mask)
        pin, keep = self._ptr_array(list(frames))
        self._shard_keep = (pin, keep, m)
        ref_buf, pred_buf = C.c_void_p(), C.c_void_p()
        ref_bytes, pred_bytes = C.c_int64(), C.c_int64()
        _capi.check(_capi.lib().vsr_sttn_shard_begin(self._h, C.cast(pin, C.POINTER(C.c_void_p)), len(frames), H, W, _capi.ptr(m, C.c_uint8), int(rank),
                                                     int(world), C.byref(ref_buf), C.byref(ref_bytes), C.byref(pred_buf), C.byref(pred_bytes)))
        return int(ref_buf.value), int(ref_bytes.value), int(pred_buf.value), int(pred_bytes.value)

    def copy_device(self, dst: int, src: int, nbytes: int) -> None:
        """device -> device on the engine's stream, completed on return"""
        _capi.check(_capi.lib().vsr_sttn_copy(self._h, C.c_void_p(dst), C.c_void_p(src), int(nbytes)))

    def shard_windows(self) -> None:
        _capi.check(_capi.lib().vsr_sttn_shard_windows(self._h))

    def shard_finish(self, frames: Sequence[np.ndarray]) -> None:
        pout, keep = self._ptr_array(list(frames))
        _capi.check(_capi.lib().vsr_sttn_shard_finish(self._h, C.cast(pout, C.POINTER(C.c_void_p))))

    @property
    def cuda_stream(self) -> int:
        return int(_capi.lib().vsr_sttn_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(_capi.lib().vsr_sttn_launch_count(self._h))

    def debug_read(self, name: str, shape) -> np.ndarray:
        out = np.empty(shape, np.float32)
        _capi.check(_capi.lib().vsr_sttn_debug_read(self._h, name.encode(), _capi.ptr(out, C.c_float), out.size))
        return out

    PROF_CLASSES = ("conv3x3", "conv3x3_residual", "qkv", "score", "softmax", "pv", "encoder", "decoder", "gather", "prepost")

    def profile(self) -> Dict[str, tuple]:
        """One eager pass of the staged chunk with CUDA events around every launch group: {class: (ms, groups)}."""
        n = len(self.PROF_CLASSES)
        ms, cnt = np.zeros(n, np.float32), np.zeros(n, np.int64)
        _capi.check(_capi.lib().vsr_sttn_profile(self._h, _capi.ptr(ms, C.c_float), _capi.ptr(cnt, C.c_int64), n))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.PROF_CLASSES)}

    def time_conv(self, T: int, n: int) -> np.ndarray:
        ms = np.zeros(n, np.float32)
        _capi.check(_capi.lib().vsr_sttn_time_conv(self._h, T, n, _capi.ptr(ms, C.c_float)))
        return ms


class _DevicePointer:
    """raw device memory as a CUDA-array-interface object (so that torch can wrap it without a copy)"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _nccl_all_gather(engine, ptr: int, region_bytes: int, rank: int, world: int) -> None:
    """All-gather of an engine exchange buffer laid out [world][region_bytes] (region r = rank r's contribution) with torch.distributed / NCCL
    over NVLink: this rank's region goes device-to-device into a torch tensor, `all_gather_into_tensor` fills the gathered tensor, and that
    comes back device-to-device (vsr_sttn_copy) — the engine's memory is never handed to another allocator, nothing touches the host."""
    import torch
    import torch.distributed as dist

    dev = torch.device("cuda", engine._dev)
    cache = engine.__dict__.setdefault("_gather_bufs", {})
    if cache.get("n") != (region_bytes, world):
        cache.update(n=(region_bytes, world), send=torch.empty(region_bytes, dtype=torch.uint8, device=dev),
                     recv=torch.empty(region_bytes * world, dtype=torch.uint8, device=dev))
    send, recv = cache["send"], cache["recv"]
    torch.cuda.synchronize(dev)
    engine.copy_device(send.data_ptr(), ptr + rank * region_bytes, region_bytes)
    dist.all_gather_into_tensor(recv, send)
    torch.cuda.synchronize(dev)
    engine.copy_device(ptr, recv.data_ptr(), region_bytes * world)


def _in_ab_sections(frame_no, ab_sections) -> bool:
    """backend/tools/inpaint_tools.py:303-321."""
    if not ab_sections:
        return True
    return any(frame_no in s for s in ab_sections)


class STTNAutoInpaint:
    """Drop-in for backend/inpaint/sttn_auto_inpaint.py:167 `STTNAutoInpaint` (whole-video driver)."""

    def __init__(self, device, model_path, video_path, mask_path=None, clip_gap=None):
        self.sttn_inpaint = STTNInpaint(device, model_path)
        self.video_path = video_path
        self.mask_path = mask_path
        self.video_out_path = os.path.join(os.path.dirname(os.path.abspath(self.video_path)),
                                           f"{os.path.basename(self.video_path).rsplit('.', 1)[0]}_no_sub.mp4")
        self.clip_gap = config.getSttnMaxLoadNum() if clip_gap is None else clip_gap  # :194-197

    def read_frame_info_from_video(self):
        import cv2

        reader = cv2.VideoCapture(self.video_path)
        info = {"W_ori": int(reader.get(cv2.CAP_PROP_FRAME_WIDTH) + 0.5), "H_ori": int(reader.get(cv2.CAP_PROP_FRAME_HEIGHT) + 0.5),
                "fps": reader.get(cv2.CAP_PROP_FPS), "len": int(reader.get(cv2.CAP_PROP_FRAME_COUNT) + 0.5)}
        return reader, info

    def __call__(self, input_mask=None, input_sub_remover=None, tbar=None):
        """sttn_auto_inpaint.py:199-336.  Chunks of clip_gap frames are independent (:242-245); frames
        outside the A/B sections pass through untouched; the writer receives every frame in order."""
        import cv2

        reader = writer = None
        try:
            reader, info = self.read_frame_info_from_video()
            if input_sub_remover is not None:
                ab_sections = input_sub_remover.ab_sections
                writer = input_sub_remover.video_writer
            else:
                ab_sections = None
                writer = cv2.VideoWriter(self.video_out_path, cv2.VideoWriter_fourcc(*"mp4v"), info["fps"],
                                         (info["W_ori"], info["H_ori"]))
            if input_mask is None:
                mask = self.sttn_inpaint.read_mask(self.mask_path)[:, :, 0] * 255
            else:
                mask = input_mask
            n, gap = info["len"], self.clip_gap
            gui = input_sub_remover is not None and getattr(input_sub_remover, "gui_mode", False)
            eng = self.sttn_inpaint
            single_strip = len(get_inpaint_area_by_mask(info["W_ori"], info["H_ori"], int(info["W_ori"] * 3 / 16),
                                                       (np.asarray(mask) > 127).astype(np.uint8))) == 1

            def finish(job):
                frames, originals, sel, ticket = job
                if ticket is not None:
                    eng.collect(ticket, [frames[i] for i in sel])
                for i, frame in enumerate(frames):
                    writer.write(frame)
                    if input_sub_remover is not None:
                        if tbar is not None:
                            input_sub_remover.update_progress(tbar, increment=1)
                        if gui:
                            input_sub_remover.update_preview_with_comp(originals[i], frame)

            pending = None  # chunk whose kernels are running while the next one is decoded and uploaded
            for start in range(0, n, gap):
                end = min(start + gap, n)
                frames = []
                for j in range(start, end):
                    ok, image = reader.read()
                    if not ok:
                        print(f"Warning: Failed to read frame {j}.")
                        break
                    frames.append(image)
                if not frames:
                    print(f"Warning: No valid frames found in range {start + 1}-{end}. Skipping this segment.")
                    continue
                originals = [f.copy() for f in frames] if gui else None
                sel = [i for i in range(len(frames)) if _in_ab_sections(start + i, ab_sections)]
                ticket = None
                if sel:
                    if single_strip:
                        ticket = eng.submit([frames[i] for i in sel], mask)
                    else:
                        eng.inpaint_inplace([frames[i] for i in sel], mask)
                if pending is not None:
                    finish(pending)
                pending = (frames, originals, sel, ticket)
            if pending is not None:
                finish(pending)
        except _capi.VsrError:
            raise  # engine / CUDA failures are never swallowed
        except Exception as e:  # the reference prints and carries on (:329-331)
            print(f"Error during video processing: {str(e)}")
        finally:
            if reader is not None:
                reader.release()
            if writer is not None:
                writer.release()
