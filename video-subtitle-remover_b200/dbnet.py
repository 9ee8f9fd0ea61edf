"""PP-OCRv5 DBNet text detector on the B200.

The reference calls third-party `paddleocr.TextDetection(model_dir=...).predict(img)` on the CPU
(backend/tools/subtitle_detect.py:43-58).  What the reference ships and pins is the model itself:
`backend/models/V5/ch_det/inference.json` (Paddle PIR program), `inference.pdiparams`, `inference.yml`.
This module compiles that program into calls on the device-tensor runtime of the C ABI (`vsr_rt_*`,
include/vsr_b200.h): batch-norm, bias and ReLU are folded into the convolutions, dense convolutions run on the
tcgen05 implicit-GEMM kernel, everything stays in NHWC fp16 on the device until the probability map.
`TextDetector.predict` returns `[{"dt_polys": ndarray[N,4,2]}]` like the paddleocr object it replaces.

Pre/post-processing follow PaddleX's DetResizeForTest / NormalizeImage / DBPostProcess as published (SURVEY
Appendix A.5 — the reference has no test or golden vector for them).
"""
import ctypes as C
import json
import os
import struct
from typing import Dict, List, Optional

import numpy as np

from . import _capi
from .sttn_auto_inpaint import _device_index

THRESH, BOX_THRESH, MAX_CANDIDATES, UNCLIP_RATIO, RESIZE_LONG = 0.3, 0.6, 1000, 1.5, 960  # inference.yml:22-53
_STORE_LIMIT, _STORE_TARGET = 4096.0, 1024.0   # fp16 storage: rescale a tensor above LIMIT so that its max lands near TARGET


# ------------------------------------------------------------------------------------------------ model files
def _read_params(path: str) -> List[np.ndarray]:
    """inference.pdiparams: records {u32, u64 lod levels, u32, i32 desc_len, TensorDesc proto, raw fp32} in
    sorted(parameter name) order."""
    buf = open(path, "rb").read()
    pos, out = 0, []

    def varint(b, j):
        v = s = 0
        while True:
            c = b[j]
            j += 1
            v |= (c & 127) << s
            s += 7
            if c < 128:
                return v, j

    while pos < len(buf):
        pos += 4
        (lod,) = struct.unpack_from("<Q", buf, pos)
        pos += 8
        for _ in range(lod):
            (n,) = struct.unpack_from("<Q", buf, pos)
            pos += 8 + n
        pos += 4
        (dlen,) = struct.unpack_from("<i", buf, pos)
        pos += 4
        desc, j, dims = buf[pos:pos + dlen], 0, []
        pos += dlen
        while j < len(desc):
            tag = desc[j]
            j += 1
            if tag & 7 == 0:
                v, j = varint(desc, j)
                if tag >> 3 != 1:
                    dims.append(v)
            else:
                ln, j = varint(desc, j)
                end = j + ln
                while j < end:
                    v, j = varint(desc, j)
                    dims.append(v)
        cnt = int(np.prod(dims)) if dims else 1
        out.append(np.frombuffer(buf, "<f4", cnt, pos).reshape(dims).copy())
        pos += 4 * cnt
    return out


class _Node:
    __slots__ = ("kind", "ins", "out", "attrs", "name")

    def __init__(self, kind, ins, out, attrs, name=None):
        self.kind, self.ins, self.out, self.attrs, self.name = kind, ins, out, attrs, name


def _load_program(model_dir: str):
    prog = json.load(open(os.path.join(model_dir, "inference.json")))
    nodes = []
    for o in prog["program"]["regions"][0]["blocks"][0]["ops"]:
        kind = o["#"].split(".", 1)[-1]
        outs = o.get("O", [])
        out = outs["%"] if isinstance(outs, dict) else (outs[0]["%"] if outs else None)
        attrs = {}
        for a in o.get("A", []):
            if isinstance(a, dict):
                t = a["AT"]
                attrs[a["N"]] = [x["D"] for x in t["D"]] if t["#"] == "0.a_array" else t.get("D")
        nodes.append(_Node(kind, [i["%"] for i in o.get("I", [])], out, attrs, o["A"][3] if kind == "p" else None))
    names = sorted(n.name for n in nodes if n.kind == "p")
    tensors = _read_params(os.path.join(model_dir, "inference.pdiparams"))
    if len(names) != len(tensors):
        raise _capi.VsrError(f"{model_dir}: {len(names)} parameters in the program, {len(tensors)} in pdiparams")
    return nodes, dict(zip(names, tensors))


def _r(x, m):
    return (x + m - 1) // m * m


class _Tensor:
    """NHWC fp16 device tensor; `c` real channels, pitch `cp`; `perm[l]` = physical channel of logical channel l.
    The buffer holds value * scale (a power of two, see `_calibrate`); `follow` ties the scale to another tensor's
    (ReLU, pooling, up-sampling and concat are positively homogeneous: they keep the scale of their input)."""

    def __init__(self, ptr, c, h, w, cp, perm=None, follow=None, n=1):
        self.ptr, self.c, self.h, self.w, self.cp, self.perm, self.n = ptr, c, h, w, cp, perm, n   # n images [n,h,w,cp]
        self._scale, self.follow = 1.0, follow

    @property
    def pixels(self):
        return self.n * self.h * self.w

    @property
    def scale(self):
        return self.follow.scale if self.follow is not None else self._scale

    @scale.setter
    def scale(self, v):
        if self.follow is not None:
            raise _capi.VsrError("the scale of this tensor is tied to its producer's input")
        self._scale = float(v)


class _Conv:
    """conv (+ folded bias / BN / ReLU): stored out = acc * (s_out / (s_in * s_w)) + bias * s_out, where the layer's weights were packed
    multiplied by the power of two s_w (see `_weight_scale`)."""
    rescalable = True

    def __init__(self, lid, x, y, relu, wscale=1.0):
        self.lid, self.x, self.y, self.relu, self.wscale = lid, x, y, relu, wscale

    def run(self, rt):
        rt.conv(self.lid, self.x, self.y, self.relu, self.y.scale / (self.x.scale * self.wscale), self.y.scale)


def _weight_scale(w: np.ndarray) -> float:
    """Power of two that lifts a layer's weights out of fp16's subnormal range.  The convs behind the un-normalised neck fold a tiny
    batch-norm scale (their inputs reach 1e5): weights of 1e-6 .. 1e-8 round to fp16 subnormals or to zero — that, not ordinary rounding,
    was the detector's parity limit (max |dp| 0.3 -> 0.06 in the fp16 simulation).  The inverse goes into the epilogue multiplier."""
    m = float(np.abs(w).max())
    if m == 0.0 or m >= 2.0 ** -6:
        return 1.0
    return float(2.0 ** -int(np.floor(np.log2(m))))



class _Add:
    """a + b (+ ReLU) of two activations that may sit at different scales."""
    rescalable = True

    def __init__(self, op, a, b, y):
        self.op, self.a, self.b, self.y = op, a, b, y

    def run(self, rt):
        rt.elementwise(self.op, self.a, self.b, self.y, alpha=self.y.scale / self.a.scale, beta=self.y.scale / self.b.scale)


class _Unary:
    """relu (op 1, output follows the input scale), sigmoid (op 3, output at scale 1), x*alpha+beta (op 6)."""
    rescalable = False

    def __init__(self, op, x, y, alpha=1.0, beta=0.0):
        self.op, self.x, self.y, self.alpha, self.beta = op, x, y, alpha, beta

    def run(self, rt):
        if self.op == 3:
            rt.elementwise(3, self.x, None, self.y, alpha=1.0 / self.x.scale)
        elif self.op == 6:
            rt.elementwise(6, self.x, None, self.y, alpha=self.alpha, beta=self.beta * self.x.scale)
        else:
            rt.elementwise(self.op, self.x, None, self.y)


class _HSwish:
    """hardswish followed by the exported learnable scalar affine: out = a * hswish(x) + c (PPLCNetV3's Act + LearnableAffineBlock).
    Not homogeneous, so the output has its own scale: stored = s_out * (a * hswish(stored_in / s_in) + c)."""
    rescalable = True

    def __init__(self, x, y, a=1.0, c=0.0):
        self.x, self.y, self.a, self.c = x, y, a, c

    def run(self, rt):
        rt.hswish_affine(self.x, self.y, 1.0 / self.x.scale, self.a * self.y.scale, self.c * self.y.scale)


class _SE:
    """squeeze-and-excitation: gate = hardsigmoid(fc2(relu(fc1(mean(x))))), y = x * gate (SEModule) or x + x * gate (RSELayer).
    Homogeneous in x for a given gate: y keeps the scale of x."""
    rescalable = False

    def __init__(self, se_id, x, y, gate, zeros):
        self.se_id, self.x, self.y, self.gate, self.zeros = se_id, x, y, gate, zeros

    def run(self, rt):
        for j in range(self.x.n):          # the gate is a per-image statistic: one image of the batch at a time (same stream, one gate buffer)
            xj, yj = _image(self.x, j), _image(self.y, j)
            rt.se_gate(self.se_id, xj, 1.0 / self.x.scale, self.gate)
            rt.elementwise(4, xj, None, yj, scale=self.gate, shift=self.zeros)


def _image(t, j):
    """image j of a batch tensor as a single-image view"""
    return t if t.n == 1 else _Tensor(t.ptr + j * t.h * t.w * t.cp * 2, t.c, t.h, t.w, t.cp, t.perm, n=1)


class _Affine:
    """per-channel x*sc + sh (+ ReLU) for a batch-norm / bias that has no convolution to fold into."""
    rescalable = False

    def __init__(self, op, x, y, sc, sh):
        self.op, self.x, self.y, self.sc, self.sh = op, x, y, sc, sh
        self._dev = None   # (scale the shift was uploaded for, scale ptr, shift ptr)

    def run(self, rt):
        if self._dev is None or self._dev[0] != self.x.scale:
            self._dev = (self.x.scale, rt.upload_f32(self.sc), rt.upload_f32(self.sh * self.x.scale))
        rt.elementwise(self.op, self.x, None, self.y, scale=self._dev[1], shift=self._dev[2])


class _Concat:
    """channel concat: every part is copied into its slice of the output, which takes the smallest scale among the
    parts (parts stored at a larger scale are multiplied down on the way)."""
    rescalable = False

    def __init__(self, parts, offsets, y, new_like):
        self.parts, self.offsets, self.y, self._new_like = parts, offsets, y, new_like
        self._tmp = {}

    def run(self, rt):
        for i, p in enumerate(self.parts):
            src = p
            if p.scale != self.y.scale:
                if i not in self._tmp:
                    self._tmp[i] = self._new_like(p)
                src = self._tmp[i]
                rt.elementwise(6, p, None, src, alpha=self.y.scale / p.scale, beta=0.0)
            rt.copy_channels(src, self.y, self.offsets[i], _r(p.c, 8))


class _Call:
    """scale-preserving single launch (nearest up-sampling, max pooling)."""
    rescalable = False

    def __init__(self, fn, *args):
        self.fn, self.args = fn, args

    def run(self, rt):
        getattr(rt, self.fn)(*self.args)


class _Compiled:
    def __init__(self):
        self.steps = []      # _Conv / _Add / _Unary / _Affine / _Concat / _Call in execution order
        self.inp = None
        self.out = None
        self.values = {}     # PIR value id -> _Tensor (diagnosis: tools/dbnet_diag.py)
        self.graph = None    # CUDA graph of `steps` at the current scales
        self.calibrated = False


class _DeviceRuntime:
    """The `vsr_rt_*` C ABI behind a small interface, so that the graph compiler can also be driven by a CPU
    stand-in in the tests (tests/fake_rt.py) and its folding / layout logic checked without a GPU."""

    def __init__(self, device):
        L = _capi.lib()
        h = C.c_void_p()
        _capi.check(L.vsr_rt_create(C.byref(h), _device_index(device)))
        self.h, self.L = h, L

    def close(self):
        if self.h:
            self.L.vsr_rt_destroy(self.h)
            self.h = None

    def alloc(self, nbytes: int) -> int:
        p = C.c_uint64()
        _capi.check(self.L.vsr_rt_alloc(self.h, int(nbytes), C.byref(p)))
        return int(p.value)

    def free(self, ptr: int) -> None:
        _capi.check(self.L.vsr_rt_free(self.h, int(ptr)))

    def upload_f32(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr, np.float32)
        p = self.alloc(arr.nbytes)
        _capi.check(self.L.vsr_rt_upload(self.h, p, arr.ctypes.data_as(C.c_void_p), arr.nbytes))
        return p

    def conv_create(self, w, bias, cout, cin, cin_pitch, kh, kw, stride, pad_t, pad_l, dil, groups, transposed) -> int:
        w = np.ascontiguousarray(w, np.float32)
        bias = np.ascontiguousarray(bias, np.float32)
        lid = C.c_int32()
        f32p = C.POINTER(C.c_float)
        _capi.check(self.L.vsr_rt_conv_create(self.h, w.ctypes.data_as(f32p), bias.ctypes.data_as(f32p), cout, cin, cin_pitch, kh, kw, stride,
                                              pad_t, pad_l, dil, groups, 1 if transposed else 0, C.byref(lid)))
        return int(lid.value)

    def conv_create_split(self, w, bias, cout, cin, cin_pitch, kh, kw, pad_t, pad_l, dil) -> int:
        w = np.ascontiguousarray(w, np.float32)
        bias = np.ascontiguousarray(bias, np.float32)
        lid = C.c_int32()
        f32p = C.POINTER(C.c_float)
        _capi.check(self.L.vsr_rt_conv_create_split(self.h, w.ctypes.data_as(f32p), bias.ctypes.data_as(f32p), cout, cin, cin_pitch, kh, kw, pad_t, pad_l, dil,
                                                    C.byref(lid)))
        return int(lid.value)

    def conv(self, lid, x, y, relu, alpha=1.0, bias_scale=1.0):
        _capi.check(self.L.vsr_rt_conv(self.h, lid, x.ptr, x.n, x.h, x.w, y.ptr, y.cp, 0, relu, alpha, bias_scale))

    def hswish_affine(self, x, y, inv_scale_in, a, c):
        _capi.check(self.L.vsr_rt_hswish_affine(self.h, x.ptr, y.ptr, x.pixels * x.cp, inv_scale_in, a, c))

    def se_create(self, w1, b1, w2, b2, slope, offset, residual) -> int:
        w1, b1, w2, b2 = (np.ascontiguousarray(v, np.float32) for v in (w1, b1, w2, b2))
        f32p, sid = C.POINTER(C.c_float), C.c_int()
        _capi.check(self.L.vsr_rt_se_create(self.h, w1.ctypes.data_as(f32p), b1.ctypes.data_as(f32p), w2.ctypes.data_as(f32p), b2.ctypes.data_as(f32p),
                                            w1.shape[1], w1.shape[0], slope, offset, 1 if residual else 0, C.byref(sid)))
        return int(sid.value)

    def se_gate(self, se_id, x, inv_scale, gate_ptr):
        _capi.check(self.L.vsr_rt_se_gate(self.h, se_id, x.ptr, x.pixels, x.cp, inv_scale, gate_ptr))

    def absmax(self, t) -> float:
        out = C.c_float()
        _capi.check(self.L.vsr_rt_absmax(self.h, t.ptr, t.pixels * t.cp, C.byref(out)))
        return float(out.value)

    def overflow(self) -> bool:
        f = C.c_int()
        _capi.check(self.L.vsr_rt_overflow(self.h, C.byref(f)))
        return bool(f.value)

    def capture_begin(self):
        _capi.check(self.L.vsr_rt_capture_begin(self.h))

    def capture_end(self) -> int:
        g = C.c_int()
        _capi.check(self.L.vsr_rt_capture_end(self.h, C.byref(g)))
        return int(g.value)

    def graph_launch(self, g):
        _capi.check(self.L.vsr_rt_graph_launch(self.h, g))

    def graph_destroy(self, g):
        _capi.check(self.L.vsr_rt_graph_destroy(self.h, g))

    def elementwise(self, op, a, b, y, scale=0, shift=0, alpha=1.0, beta=1.0):
        _capi.check(self.L.vsr_rt_elementwise(self.h, op, a.ptr, b.ptr if b is not None else 0, y.ptr, a.pixels * a.cp, a.cp, scale, shift,
                                              alpha, beta))

    def upsample(self, x, y, s):
        _capi.check(self.L.vsr_rt_upsample_nearest(self.h, x.ptr, x.n, x.h, x.w, x.cp, s, y.ptr, y.cp, 0))

    def maxpool(self, x, y):
        _capi.check(self.L.vsr_rt_maxpool2x2s1(self.h, x.ptr, x.n, x.h, x.w, x.cp, y.ptr))

    def copy_channels(self, src, dst, dst_off, channels):
        _capi.check(self.L.vsr_rt_copy_channels(self.h, src.ptr, src.cp, dst.ptr, dst.cp, dst_off, channels, src.pixels))

    def preprocess(self, img, inp, rh, rw, slot=0):
        _capi.check(self.L.vsr_rt_det_preprocess(self.h, _capi.ptr(img, C.c_uint8), img.shape[0], img.shape[1], inp.ptr + slot * rh * rw * inp.cp * 2, rh, rw,
                                                 inp.cp))

    def download(self, t) -> np.ndarray:
        host = np.empty((t.h, t.w, t.cp), np.float16)
        _capi.check(self.L.vsr_rt_download(self.h, t.ptr, host.ctypes.data_as(C.c_void_p), host.nbytes))
        return host

    def sync(self):
        _capi.check(self.L.vsr_rt_sync(self.h))

    def download_channel(self, t, ch, slot=0) -> np.ndarray:
        """channel `ch` of image `slot` of a [n,h,w,cp] tensor as fp32 [h,w], divided by the tensor scale"""
        host = np.empty((t.h, t.w), np.float32)
        _capi.check(self.L.vsr_rt_download_channel(self.h, t.ptr + slot * t.h * t.w * t.cp * 2, t.h * t.w, t.cp, ch, 1.0 / t.scale,
                                                   host.ctypes.data_as(C.POINTER(C.c_float))))
        return host

    @property
    def launch_count(self) -> int:
        return int(self.L.vsr_rt_launch_count(self.h))


class TextDetector:
    """Stand-in for `paddleocr.TextDetection(model_name, model_dir, device=...)` (subtitle_detect.py:47-52)."""

    def __init__(self, model_dir: str, device="cuda:0", model_name: Optional[str] = None, runtime=None, precise_weights: Optional[bool] = None):
        """precise_weights: keep the weights of the dense stride-1 convs (up to 40 taps) as hi + lo fp16 halves.  The detector's parity is
        limited by WEIGHT rounding in its head (fp32 weights + fp16 activations: max |dp| 0.03 instead of 0.34 in the fp16 simulation,
        profiles/fp16_forecast_r1.md); the path has not run on a B200 yet, so it is opt-in (argument or VSR_DET_PRECISE_WEIGHTS=1)."""
        self.precise_weights = (os.environ.get("VSR_DET_PRECISE_WEIGHTS") == "1") if precise_weights is None else bool(precise_weights)
        self.model_dir, self.model_name = model_dir, model_name
        self._nodes, self._params = _load_program(model_dir)
        self._rt = runtime if runtime is not None else _DeviceRuntime(device)
        self._programs: Dict[tuple, _Compiled] = {}

    def __del__(self):
        rt = getattr(self, "_rt", None)
        if rt is not None:
            try:
                rt.close()
            except Exception:
                pass
            self._rt = None

    # -------------------------------------------------------------------------------------------- runtime helpers
    def _new(self, c, h, w, perm=None, follow=None) -> _Tensor:
        cp = _r(max(c, 1), 64)
        n = getattr(self, "_n", 1)          # images per launch of the program being compiled
        return _Tensor(self._rt.alloc(n * h * w * cp * 2), c, h, w, cp, perm, follow, n=n)

    def _new_like(self, t: _Tensor) -> _Tensor:
        return _Tensor(self._rt.alloc(t.pixels * t.cp * 2), t.c, t.h, t.w, t.cp, t.perm, n=t.n)

    # -------------------------------------------------------------------------------------------- graph compiler
    def _compile(self, H: int, W: int, N: int = 1) -> _Compiled:
        self._n = N
        rt = self._rt
        nodes, params = self._nodes, self._params
        prod = {n.out: n for n in nodes if n.out is not None}
        users: Dict[int, List[_Node]] = {}
        for n in nodes:
            for i in n.ins:
                users.setdefault(i, []).append(n)
        val: Dict[int, object] = {}   # value id -> _Tensor | ndarray | python constant
        done = set()
        prog = _Compiled()
        # constants first: the program lists `reshape(param)` bias operands AFTER the conv they are added to
        for n in nodes:
            if n.kind == "p":
                val[n.out] = params[n.name]
            elif n.kind in ("full_int_array", "full"):
                val[n.out] = n.attrs["value"]
            elif n.kind == "reshape" and isinstance(val.get(n.ins[0]), np.ndarray) and n.ins[1] in val:
                val[n.out] = np.asarray(val[n.ins[0]]).reshape([int(d) for d in val[n.ins[1]]])
            else:
                continue
            done.add(id(n))

        def single_user(v, kind):
            u = users.get(v, [])
            return u[0] if len(u) == 1 and u[0].kind == kind else None

        def const_of(v):
            x = val.get(v)
            return x if x is not None and not isinstance(x, _Tensor) else None

        def bias_operand(add_node, act_id):
            """`add(x, reshape(param))`: the per-channel bias vector, else None."""
            other = [i for i in add_node.ins if i != act_id]
            if len(other) != 1:
                return None
            c = const_of(other[0])
            return np.asarray(c, np.float32).reshape(-1) if isinstance(c, np.ndarray) else None

        def scalar_of(v):
            """value of a one-element constant (a learnable scalar of the exported blocks), else None"""
            c = const_of(v)
            if isinstance(c, np.ndarray) and c.size == 1:
                return float(c.reshape(-1)[0])
            return None

        def scalar_affine(cur):
            """`multiply(scalar, x)` then `add(., scalar)` hanging off value `cur` as single consumers: (a, c, nodes, out id)."""
            a, c, used = 1.0, 0.0, []
            u = single_user(cur, "multiply")
            if u is not None:
                k = [scalar_of(i) for i in u.ins if i != cur]
                if len(k) == 1 and k[0] is not None:
                    a, cur = k[0], u.out
                    used.append(u)
            u = single_user(cur, "add")
            if u is not None:
                k = [scalar_of(i) for i in u.ins if i != cur]
                if len(k) == 1 and k[0] is not None:
                    c, cur = k[0], u.out
                    used.append(u)
            return a, c, used, cur

        def emit_conv(n: _Node):
            x: _Tensor = val[n.ins[0]]
            w = np.asarray(val[n.ins[1]], np.float32)
            a = n.attrs
            transposed = n.kind == "conv2d_transpose"
            groups = int(a.get("groups", 1))
            cout = w.shape[1] if transposed else w.shape[0]
            kh, kw = int(w.shape[2]), int(w.shape[3])
            stride, dil = int(a["strides"][0]), int(a["dilations"][0])
            if a.get("padding_algorithm") == "SAME":  # out = ceil(in/stride); extra pixel on the bottom / right
                tot_h = max((-(-x.h // stride) - 1) * stride + (kh - 1) * dil + 1 - x.h, 0)
                tot_w = max((-(-x.w // stride) - 1) * stride + (kw - 1) * dil + 1 - x.w, 0)
                pad_t, pad_l = tot_h // 2, tot_w // 2
            else:
                pad_t, pad_l = int(a["paddings"][0]), int(a["paddings"][1])
            bias = np.zeros(cout, np.float32)
            w = w.copy()
            cur, relu = n.out, 0
            # fold: [+ bias add] [-> batch_norm] [-> relu], each only when it is the single consumer
            u = single_user(cur, "add")
            if u is not None:
                b = bias_operand(u, cur)
                if b is not None and b.size == cout:
                    bias += b
                    done.add(id(u))
                    cur = u.out
            u = single_user(cur, "batch_norm_")
            if u is not None and u.ins[0] == cur:
                mean, var, gamma, beta = (np.asarray(val[i], np.float32) for i in u.ins[1:5])
                s = gamma / np.sqrt(var + np.float32(u.attrs["epsilon"]))
                if transposed:
                    w *= s[None, :, None, None]
                else:
                    w *= s[:, None, None, None]
                bias = (bias - mean) * s + beta
                done.add(id(u))
                cur = u.out
            a_s, c_s, used, nxt = scalar_affine(cur)     # LearnableAffineBlock after the (re-parameterised) conv
            if used:
                w *= np.float32(a_s)
                bias = bias * np.float32(a_s) + np.float32(c_s)
                done.update(id(m) for m in used)
                cur = nxt
            u = single_user(cur, "relu")
            if u is not None:
                relu = 1
                done.add(id(u))
                cur = u.out
            if groups == 1 and not transposed and cout >= 8 and cout % 8:   # the tensor-core conv stores 8-channel groups
                pad = _r(cout, 8) - cout
                w = np.concatenate([w, np.zeros((pad,) + w.shape[1:], np.float32)])
                bias = np.concatenate([bias, np.zeros(pad, np.float32)])
            cin_l = w.shape[0] if transposed else w.shape[1] * (groups if groups > 1 else 1)
            if groups == 1 and x.perm is not None:  # logical input channel l lives at physical channel perm[l]
                wp = np.zeros((w.shape[0], x.cp) + w.shape[2:], np.float32) if not transposed else None
                if transposed:
                    raise _capi.VsrError("permuted input into a transposed conv is not supported")
                wp[:, x.perm] = w
                w, cin_eff = wp, x.cp
            else:
                cin_eff = x.c if groups == 1 else x.c
            if stride == 2 and not transposed:
                oh, ow = (x.h + 2 * pad_t - (kh - 1) * dil - 1) // 2 + 1, (x.w + 2 * pad_l - (kw - 1) * dil - 1) // 2 + 1
            elif transposed:
                oh, ow = 2 * x.h, 2 * x.w
            else:
                oh, ow = x.h, x.w
            y = self._new(cout, oh, ow)
            wscale = _weight_scale(w) if (groups == 1 and not transposed and cin_eff >= 16 and bias.size >= 8) else 1.0   # tensor-core layers hold fp16 weights
            w = w * np.float32(wscale)
            if self.precise_weights and groups == 1 and not transposed and stride == 1 and cin_eff >= 16 and bias.size >= 8 and 2 * kh * kw <= 81:
                lid = rt.conv_create_split(w, bias, int(bias.size), int(cin_eff), x.cp, kh, kw, pad_t, pad_l, dil)
            else:
                lid = rt.conv_create(w, bias, int(bias.size), int(cin_eff), x.cp, kh, kw, stride, pad_t, pad_l, dil, groups, transposed)
            prog.steps.append(_Conv(lid, x, y, relu, wscale))
            val[cur] = y

        def channel_vectors(x: _Tensor, mul, add):
            sc, sh = np.zeros(x.cp, np.float32), np.zeros(x.cp, np.float32)
            idx = x.perm if x.perm is not None else np.arange(x.c)
            sc[idx], sh[idx] = mul, add
            return sc, sh

        for n in nodes:
            if id(n) in done:
                continue
            k = n.kind
            if k == "data":
                prog.inp = self._new(3, H, W)
                val[n.out] = prog.inp
            elif k == "reshape":
                raise _capi.VsrError("reshape of an activation is not supported")
            elif k == "combine":
                val[n.out] = [val[i] for i in n.ins]
            elif k in ("conv2d", "depthwise_conv2d", "conv2d_transpose"):
                emit_conv(n)
            elif k == "batch_norm_":  # not preceded by a conv: per-channel affine (+ relu)
                x = val[n.ins[0]]
                mean, var, gamma, beta = (np.asarray(val[i], np.float32) for i in n.ins[1:5])
                g = gamma / np.sqrt(var + np.float32(n.attrs["epsilon"]))
                sc, sh = channel_vectors(x, g, beta - mean * g)
                cur, op = n.out, 4
                u = single_user(cur, "relu")
                if u is not None:
                    op, cur = 5, u.out
                    done.add(id(u))
                val[cur] = self._new(x.c, x.h, x.w, x.perm, follow=x)
                prog.steps.append(_Affine(op, x, val[cur], sc, sh))
            elif k == "relu":
                x = val[n.ins[0]]
                val[n.out] = self._new(x.c, x.h, x.w, x.perm, follow=x)
                prog.steps.append(_Unary(1, x, val[n.out]))
            elif k == "hardswish":
                x = val[n.ins[0]]
                a_s, c_s, used, cur = scalar_affine(n.out)
                done.update(id(m) for m in used)
                val[cur] = self._new(x.c, x.h, x.w, x.perm)
                prog.steps.append(_HSwish(x, val[cur], a_s, c_s))
            elif k == "sigmoid":
                x = val[n.ins[0]]
                val[n.out] = self._new(x.c, x.h, x.w, x.perm)
                prog.steps.append(_Unary(3, x, val[n.out]))
            elif k == "add":
                a, b = val[n.ins[0]], val[n.ins[1]]
                if isinstance(a, _Tensor) and isinstance(b, _Tensor):
                    if (a.c, a.h, a.w) != (b.c, b.h, b.w) or a.perm is not None or b.perm is not None:
                        raise _capi.VsrError("add of mismatching activations")
                    cur, op = n.out, 0
                    u = single_user(cur, "relu")
                    if u is not None:
                        op, cur = 2, u.out
                        done.add(id(u))
                    val[cur] = self._new(a.c, a.h, a.w)
                    prog.steps.append(_Add(op, a, b, val[cur]))
                else:  # activation + per-channel constant that was not folded into a conv
                    x, cst = (a, b) if isinstance(a, _Tensor) else (b, a)
                    cst = np.asarray(cst, np.float32).reshape(-1)
                    sc, sh = channel_vectors(x, 1.0, cst if cst.size == x.c else float(cst[0]))
                    val[n.out] = self._new(x.c, x.h, x.w, x.perm, follow=x)
                    prog.steps.append(_Affine(4, x, val[n.out], sc, sh))
            elif k == "scale":
                x = val[n.ins[0]]
                val[n.out] = self._new(x.c, x.h, x.w, x.perm, follow=x)
                prog.steps.append(_Unary(6, x, val[n.out], alpha=float(np.asarray(val[n.ins[1]]).reshape(-1)[0]),
                                         beta=float(n.attrs.get("bias", 0.0))))
            elif k == "concat":
                parts: List[_Tensor] = val[n.ins[0]]
                if int(np.asarray(val[n.ins[1]]).reshape(-1)[0]) != 1:
                    raise _capi.VsrError("only channel concat is supported")
                total = sum(p.c for p in parts)
                # physical layout: 8-aligned parts first, ragged parts (the 1-channel probability map) last
                order = sorted(range(len(parts)), key=lambda i: (parts[i].c % 8 != 0, i))
                perm, off, phys = np.zeros(total, np.int64), 0, {}
                for i in order:
                    phys[i] = off
                    off += _r(parts[i].c, 8)
                lo = 0
                for i, p in enumerate(parts):
                    src = p.perm if p.perm is not None else np.arange(p.c)
                    perm[lo:lo + p.c] = phys[i] + src
                    lo += p.c
                y = _Tensor(rt.alloc(parts[0].pixels * _r(off, 64) * 2), total, parts[0].h, parts[0].w, _r(off, 64),
                            None if np.array_equal(perm, np.arange(total)) else perm, n=parts[0].n)
                prog.steps.append(_Concat(parts, [phys[i] for i in range(len(parts))], y, self._new_like))
                val[n.out] = y
            elif k == "nearest_interp":
                x = val[n.ins[0]]
                s = int(round(float((n.attrs.get("scale") or [2.0])[0])))
                y = _Tensor(rt.alloc(x.pixels * s * s * x.cp * 2), x.c, x.h * s, x.w * s, x.cp, x.perm, follow=x, n=x.n)
                prog.steps.append(_Call("upsample", x, y, s))
                val[n.out] = y
            elif k == "pool2d" and n.attrs.get("adaptive") and n.attrs.get("pooling_type") == "avg":
                # squeeze-and-excitation: pool(1x1) -> conv1x1 + bias -> relu -> conv1x1 + bias -> hardsigmoid -> x * gate [-> x + .]
                x = val[n.ins[0]]
                if [int(v) for v in val[n.ins[1]]] != [1, 1] or x.perm is not None:
                    raise _capi.VsrError("only global average pooling of a plain tensor is supported")

                def conv_bias(node):
                    wt = np.asarray(val[node.ins[1]], np.float32)
                    u = single_user(node.out, "add")
                    b = bias_operand(u, node.out) if u is not None else None
                    if wt.shape[2:] != (1, 1) or b is None:
                        raise _capi.VsrError("unexpected squeeze-and-excitation layout")
                    return wt[:, :, 0, 0], b, u

                c1 = single_user(n.out, "conv2d")
                w1, b1, a1 = conv_bias(c1)
                r = single_user(a1.out, "relu")
                c2 = single_user(r.out, "conv2d") if r is not None else None
                if c2 is None:
                    raise _capi.VsrError("unexpected squeeze-and-excitation layout")
                w2, b2, a2 = conv_bias(c2)
                hs = single_user(a2.out, "hardsigmoid")
                mul = single_user(hs.out, "multiply") if hs is not None else None
                if mul is None or n.ins[0] not in mul.ins:
                    raise _capi.VsrError("unexpected squeeze-and-excitation layout")
                cur, residual = mul.out, False
                u = single_user(cur, "add")
                if u is not None and n.ins[0] in u.ins:          # RSELayer: x + x * gate
                    cur, residual = u.out, True
                    done.add(id(u))
                done.update(id(m) for m in (c1, a1, r, c2, a2, hs, mul))
                se_id = rt.se_create(w1, b1, w2, b2, float(hs.attrs["slope"]), float(hs.attrs["offset"]), residual)
                y = self._new(x.c, x.h, x.w, follow=x)
                prog.steps.append(_SE(se_id, x, y, rt.upload_f32(np.zeros(x.cp, np.float32)), rt.upload_f32(np.zeros(x.cp, np.float32))))
                val[cur] = y
            elif k == "pool2d":
                x = val[n.ins[0]]
                ks = [int(v) for v in val[n.ins[1]]]
                a = n.attrs
                if not (a["pooling_type"] == "max" and ks == [2, 2] and a["strides"] == [1, 1] and a["padding_algorithm"] == "SAME" and not a.get("adaptive")):
                    raise _capi.VsrError(f"unsupported pool2d {a} (the mobile detector's SE blocks are not compiled yet)")
                y = self._new(x.c, x.h, x.w, x.perm, follow=x)
                prog.steps.append(_Call("maxpool", x, y))
                val[n.out] = y
            elif k == "fetch":
                prog.out = val[n.ins[0]]
            else:
                raise _capi.VsrError(f"PIR op '{k}' is not supported by the B200 detector")
        if prog.inp is None or prog.out is None:
            raise _capi.VsrError("program without data/fetch")
        prog.values = {k: t for k, t in val.items() if isinstance(t, _Tensor)}
        return prog

    # -------------------------------------------------------------------------------------------- inference
    @staticmethod
    def resize_shape(h: int, w: int):
        """DetResizeForTest(resize_long=960): long side <= 960, both sides rounded to multiples of 32."""
        ratio = RESIZE_LONG / max(h, w) if max(h, w) > RESIZE_LONG else 1.0
        rh, rw = int(h * ratio), int(w * ratio)
        return max(int(round(rh / 32) * 32), 32), max(int(round(rw / 32) * 32), 32)

    def probability_map(self, img_bgr: np.ndarray) -> np.ndarray:
        """BGR uint8 image -> DB probability map [rh, rw] fp32 (device: resize, normalise, network)."""
        img = np.ascontiguousarray(img_bgr, np.uint8)
        if img.ndim != 3 or img.shape[2] != 3:
            raise ValueError("expected a BGR uint8 image [H,W,3]")
        rh, rw = self.resize_shape(img.shape[0], img.shape[1])
        prog = self._programs.get((rh, rw))
        if prog is None:
            prog = self._programs[(rh, rw)] = self._compile(rh, rw)
        rt = self._rt
        rt.preprocess(img, prog.inp, rh, rw)
        if not prog.calibrated:
            self._calibrate(prog)        # runs the network once, layer by layer
        else:
            rt.graph_launch(prog.graph)
        out = prog.out
        ch = int(out.perm[0]) if out.perm is not None else 0
        host = rt.download_channel(out, ch)
        if rt.overflow():                # a frame whose activations outgrew the calibrated scales: shrink them and redo
            self._calibrate(prog)
            host = rt.download_channel(out, ch)
            if rt.overflow():
                raise _capi.VsrError("detector activations overflow fp16 even after rescaling")
        return host

    def probability_maps(self, images) -> List[np.ndarray]:
        """Several same-size frames through ONE launch of the network (the sampled frames of the detection pass, subtitle_detect.py:96-110):
        the detector's layers are small, so a launch per frame is latency-bound.  NOT yet run on a B200 with more than one image per launch
        (the direct / depthwise / transposed-conv kernels and the pooling kernels have only been exercised with one): opt-in."""
        imgs = [np.ascontiguousarray(i, np.uint8) for i in images]
        if not imgs or any(i.shape != imgs[0].shape or i.ndim != 3 or i.shape[2] != 3 for i in imgs):
            raise ValueError("expected same-size BGR uint8 images [H,W,3]")
        if len(imgs) == 1:
            return [self.probability_map(imgs[0])]
        N = len(imgs)
        rh, rw = self.resize_shape(imgs[0].shape[0], imgs[0].shape[1])
        prog = self._programs.get((N, rh, rw))
        if prog is None:
            prog = self._programs[(N, rh, rw)] = self._compile(rh, rw, N)
        rt = self._rt
        for j, img in enumerate(imgs):
            rt.preprocess(img, prog.inp, rh, rw, j)
        if not prog.calibrated:
            self._calibrate(prog)
        else:
            rt.graph_launch(prog.graph)
        out = prog.out
        ch = int(out.perm[0]) if out.perm is not None else 0
        maps = [rt.download_channel(out, ch, j) for j in range(N)]
        if rt.overflow():
            self._calibrate(prog)
            maps = [rt.download_channel(out, ch, j) for j in range(N)]
            if rt.overflow():
                raise _capi.VsrError("detector activations overflow fp16 even after rescaling")
        return maps

    def predict_batch(self, images):
        """`predict` for several same-size frames in one launch (see `probability_maps`)."""
        return [{"dt_polys": db_postprocess(p, i.shape[0], i.shape[1]), "prob_shape": p.shape} for p, i in zip(self.probability_maps(images), images)]

    def time_network(self, iters: int = 20) -> float:
        """ms per replay of the recorded network graph of the last-used resolution (input already on the device)."""
        import time

        prog = next(reversed(self._programs.values()))
        if prog.graph is None:
            raise _capi.VsrError("run probability_map once before timing")
        rt = self._rt
        rt.graph_launch(prog.graph)
        rt.sync()
        t0 = time.perf_counter()
        for _ in range(iters):
            rt.graph_launch(prog.graph)
        rt.sync()
        return (time.perf_counter() - t0) / iters * 1e3

    def _calibrate(self, prog: _Compiled):
        """Choose the per-tensor scales on the frame that is in `prog.inp` and record the CUDA graph.

        The detector's neck has no normalisation and reaches |x| ~ 1e5 in fp32, beyond fp16.  Everything between the
        stem and the final sigmoid is positively homogeneous (conv, ReLU, add, max-pool, nearest up-sampling, concat),
        so a tensor may be stored multiplied by a power of two without changing the result: convolutions and adds pick
        the scale of their output (a multiplier in their epilogue), everything else passes it through and the sigmoid
        divides it out.  Layers run in order; a layer whose stored output exceeds _STORE_LIMIT gets a smaller scale and
        runs again, so every later layer sees finite inputs.  Scales only ever shrink; the limit leaves 16x headroom for
        other frames, and the scaled epilogues raise a flag when a later frame still overflows (-> recalibration)."""
        rt = self._rt
        if prog.graph is not None:
            rt.graph_destroy(prog.graph)
            prog.graph = None
        for st in prog.steps:
            if isinstance(st, _Add):
                st.y.scale = min(st.y.scale, st.a.scale, st.b.scale)
            elif isinstance(st, _Concat):
                st.y.scale = min([st.y.scale] + [p.scale for p in st.parts])
            st.run(rt)
            if not st.rescalable:
                if rt.overflow():
                    raise _capi.VsrError(f"detector calibration: {type(st).__name__} step overflows fp16")
                continue
            for _ in range(12):
                m = rt.absmax(st.y)
                if m <= _STORE_LIMIT:
                    break
                st.y.scale = st.y.scale * (2.0 ** -8 if not np.isfinite(m) else 2.0 ** -int(np.ceil(np.log2(m / _STORE_TARGET))))
                st.run(rt)
            else:
                raise _capi.VsrError("detector calibration did not converge (non-finite input?)")
            rt.overflow()   # clear what the discarded attempts raised
        rt.capture_begin()
        try:
            for st in prog.steps:
                st.run(rt)
        finally:
            prog.graph = rt.capture_end()
        prog.calibrated = True

    @property
    def launch_count(self) -> int:
        return self._rt.launch_count

    def predict(self, img_bgr: np.ndarray):
        """Like paddleocr's TextDetection.predict: one result dict per image with `dt_polys` [N,4,2] int16."""
        prob = self.probability_map(img_bgr)
        return [{"dt_polys": db_postprocess(prob, img_bgr.shape[0], img_bgr.shape[1]), "prob_shape": prob.shape}]


def _order_quad(pts):
    p = sorted(np.asarray(pts).tolist(), key=lambda q: q[0])
    a, d = (0, 1) if p[1][1] > p[0][1] else (1, 0)
    b, c = (2, 3) if p[3][1] > p[2][1] else (3, 2)
    return np.array([p[a], p[b], p[c], p[d]], np.float32)


def db_postprocess(prob: np.ndarray, src_h: int, src_w: int) -> np.ndarray:
    """DBPostProcess (thresh 0.3, box_thresh 0.6, unclip 1.5, quad boxes, fast score; inference.yml:48-53) with
    OpenCV, the library PaddleOCR itself uses for it; the Clipper offset of the min-area rectangle is computed as
    that rectangle grown by area*ratio/perimeter on every side."""
    import cv2

    H, W = prob.shape
    contours, _ = cv2.findContours((prob > THRESH).astype(np.uint8) * 255, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    quads = []
    for cnt in contours[:MAX_CANDIDATES]:
        rect = cv2.minAreaRect(cnt)
        if min(rect[1]) < 3:
            continue
        q = _order_quad(cv2.boxPoints(rect))
        x0, x1 = int(np.clip(np.floor(q[:, 0].min()), 0, W - 1)), int(np.clip(np.ceil(q[:, 0].max()), 0, W - 1))
        y0, y1 = int(np.clip(np.floor(q[:, 1].min()), 0, H - 1)), int(np.clip(np.ceil(q[:, 1].max()), 0, H - 1))
        m = np.zeros((y1 - y0 + 1, x1 - x0 + 1), np.uint8)
        cv2.fillPoly(m, [(q - np.array([x0, y0], np.float32)).astype(np.int32)], 1)
        if cv2.mean(prob[y0:y1 + 1, x0:x1 + 1], m)[0] < BOX_THRESH:
            continue
        (cx, cy), (rw, rh), ang = rect
        d = rw * rh * UNCLIP_RATIO / (2 * (rw + rh))
        if min(rw, rh) + 2 * d < 5:
            continue
        g = _order_quad(cv2.boxPoints(((cx, cy), (rw + 2 * d, rh + 2 * d), ang)))
        g[:, 0] = np.clip(np.round(g[:, 0] / W * src_w), 0, src_w)
        g[:, 1] = np.clip(np.round(g[:, 1] / H * src_h), 0, src_h)
        quads.append(g.astype(np.int16))
    return np.array(quads, np.int16).reshape(-1, 4, 2)
