"""The CPU stand-in of the runtime with the ProPainter operators replaced by the REAL kernels (test infrastructure, never shipped).

`HybridRuntime` is `FakeRuntime(fp16=True)` — convolutions, the correlation GEMM and the element-wise operators of the validated runtime stay
numpy — but every operator of csrc/pp_ops.cuh (all 30 kernels) is executed from its real source: the buffers the call touches are turned into the fp16 / fp32 /
u8 / int32 images the device would hold, the product's wrapper method calls the host build of the `vsr_rt_*` entry point (tests/emu/), and
the images are written back.  Driving `PropainterInpaint` on it runs the kernels at the pipeline's own shapes, pitches, channel-slice views
and index lists, with fp16 storage between the operators: the closest thing to the device available without one."""
import numpy as np

from fake_rt import FakeRuntime
from pp_op_cases import bind_wrapper

# argument kinds per operator: h = fp16 tensor object (h? may be None), p16 = bare pointer into an fp16 buffer, f32 = bare pointer into an fp32
# buffer (f32? may be 0), f32c = small fp32 parameter vector (upload_f32), u8 / i32 = byte / index buffers, levels = correlation pyramid, s = scalar
_SIG = {
    "frames": ("s", "h"),
    "instnorm": ("h", "h", "s"),
    "context_split": ("h", "h", "h"),
    "corr_pool": ("p16", "s", "s", "s", "s", "p16", "s"),
    "corr_lookup": ("levels", "f32", "s", "s", "s", "h"),
    "gru_rh": ("h", "h", "h"),
    "gru_update": ("h", "h", "h"),
    "flow_update": ("f32", "h?", "h", "h?", "h?", "s", "s"),
    "convex_upsample": ("f32", "h", "s", "s", "s", "f32"),
    "img_prop_step": ("h", "h", "f32", "f32", "h"),
    "prop_state": ("h", "u8", "h?", "h"),
    "rfc_input": ("f32", "u8", "s", "s", "s", "s", "h"),
    "pad_replicate": ("h", "h", "s", "s"),
    "leaky": ("h", "s"),
    "temporal_taps": ("h", "h"),
    "deform_cols": ("h", "s", "h?", "s", "s", "h", "s", "f32?", "h"),
    "rfc_combine": ("h", "f32", "u8", "s", "s", "s", "s", "f32"),
    "gen_input": ("h", "u8", "i32", "s", "h"),
    "flow_down4": ("f32", "i32", "s", "s", "s", "f32"),
    "prop_masks": ("h", "h"),
    "featprop_cond": ("h", "h", "f32", "f32", "h", "h"),
    "write_extra": ("h", "h", "s", "s"),
    "unfold7s3": ("h", "h", "s"),
    "fold7s3": ("h", "h", "s", "s"),
    "layernorm": ("h", "f32c", "f32c", "h"),
    "pool4": ("h", "f32c", "f32c", "h"),
    "window_attention": ("h", "h", "h", "h", "h", "i32", "s", "i32", "s", "i32", "h"),
    "pred_to_rgb8": ("h",),
}


class _Images:
    """device images of the stand-in's buffers for the duration of one operator call: one image per buffer, so that two views of one
    buffer (channel slices, in-place operators) see each other's bytes like on the device"""

    def __init__(self, rt):
        self.rt, self.made = rt, {}

    def _image(self, ptr, kind):
        import bisect

        base = self.rt._bases[bisect.bisect_right(self.rt._bases, ptr) - 1]
        buf = self.rt.bufs[base]
        if base not in self.made:
            if kind == "half":
                img = buf.astype(np.float16)                       # one slot per fp16 element
            elif kind == "f32":
                img = np.ascontiguousarray(buf[0::2])              # one value per two slots (byte offsets stay valid)
            elif kind == "f32c":
                img = np.ascontiguousarray(buf, np.float32)
            else:
                img = buf.astype(np.uint8 if kind == "u8" else np.int32)
            self.made[base] = (kind, img)
        got, img = self.made[base]
        assert got == kind, f"buffer used as {got} and as {kind} by one operator"
        return img.ctypes.data + (ptr - base)

    def tensor(self, t):
        from vsr_b200.dbnet import _Tensor

        assert t.ptr % 16 == 0 and t.cp % 8 == 0, "16-byte accesses would fault on the device"
        return _Tensor(self._image(t.ptr, "half"), t.c, t.h, t.w, t.cp, n=getattr(t, "n", 1))

    def pointer(self, ptr, kind):
        return self._image(ptr, kind)

    def write_back(self):
        for base, (kind, img) in self.made.items():
            buf = self.rt.bufs[base]
            if kind == "half":
                if not np.isfinite(img).all():
                    self.rt._flag = True
                buf[:] = img.astype(np.float32)
            elif kind == "f32":
                buf[0::2] = img
            # parameter vectors, masks and index lists are read-only


# `on_numpy` names operators that stay on the numpy transcription (none by default: the fibers of tests/emu/cuda_emu.h make the kernels with
# warp shuffles / __syncthreads — instance norm, layer norm, window attention — affordable at pipeline sizes; with -DEMU_OS_THREADS they are not).
_LOCKSTEP = ("instnorm", "layernorm", "window_attention")


class HybridRuntime(FakeRuntime):
    def __init__(self, lib, on_numpy=(), **kw):
        super().__init__(fp16=True, **kw)
        self.real = bind_wrapper(lib)
        self.lib = lib
        self.on_numpy = set(on_numpy)
        self.real_calls = {}

    def _run_real(self, op, args):
        sig = _SIG[op]
        assert len(args) == len(sig), (op, len(args), len(sig))
        im = _Images(self)
        real_args = []
        for a, kind in zip(args, sig):
            if kind == "s":
                real_args.append(a)
            elif kind in ("h", "h?"):
                real_args.append(None if a is None else im.tensor(a))
            elif kind == "p16":
                real_args.append(im.pointer(a, "half"))
            elif kind in ("f32", "f32?"):
                real_args.append(0 if not a else im.pointer(a, "f32"))
            elif kind == "levels":
                real_args.append([(im.pointer(p, "half"),) + tuple(rest) for p, *rest in a])
            else:
                real_args.append(im.pointer(a, kind))
        before = self.lib.emu_launches(self.real.h)
        out = getattr(self.real, op)(*real_args)
        self.launches += self.lib.emu_launches(self.real.h) - before
        im.write_back()
        self.real_calls[op] = self.real_calls.get(op, 0) + 1
        return out


def _route(op):
    def method(self, *args, **kw):
        if self._rec is not None or op in self.on_numpy:   # graph capture: the stand-in records (or refuses) the call; replay comes back here
            return getattr(FakeRuntime, op)(self, *args, **kw)
        assert not kw or op == "unfold7s3", (op, kw)
        if op == "unfold7s3" and kw:
            args = args + (kw["gelu"],)
        if op == "unfold7s3" and len(args) == 2:
            args = args + (False,)
        return self._run_real(op, args)

    method.__name__ = op
    return method


for _op in _SIG:
    setattr(HybridRuntime, _op, _route(_op))
