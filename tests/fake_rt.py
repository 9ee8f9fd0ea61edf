"""CPU stand-in for the `vsr_rt_*` device runtime (test infrastructure, never shipped): implements the interface
of `vsr_b200.dbnet._DeviceRuntime` with fp32 numpy/torch on NHWC buffers that have the SAME pitches, channel
permutations and folded weights the graph compiler hands to the C ABI.  It lets `-m "not gpu"` tests check the
compiler (constant folding, BN/bias/ReLU fusion, concat layout, padding rules) against the oracle's interpreter
of the same program; the kernels themselves are checked on the GPU (tests/test_gpu_dbnet.py)."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import dbnet_oracle as D


def _r8(x):
    return (x + 7) // 8 * 8


_NOT_OPS = {"copy_bytes", "conv_create_split", "alloc", "upload_f32", "upload_bytes", "upload_to", "upload_ints", "download", "download_channel", "download_f32", "capture_begin",
            "capture_end", "graph_launch", "graph_destroy", "conv_create", "se_create", "overflow", "absmax", "close", "sync", "zero", "launch_count"}


def _fp16_storage(cls):
    """fp16=True mode: after every operator, round every tensor argument to fp16 (what the device buffers hold); fp32 side buffers
    (flow state, FFT staging, residual masters) are addressed by raw pointers and stay fp32, like on the device."""
    def wrap(fn):
        def inner(self, *args, **kw):
            for a in list(args) + list(kw.values()):       # every device kernel moves 8 halves (16 bytes) at a time on its tensor arguments
                if hasattr(a, "ptr") and hasattr(a, "cp"):
                    assert a.ptr % 16 == 0 and a.cp % 8 == 0, f"{fn.__name__}: tensor at +{a.ptr % 16} bytes, pitch {a.cp}: 16-byte accesses would fault on the device"
            out = fn(self, *args, **kw)
            if self.fp16 and self._rec is None:
                for a in list(args) + list(kw.values()):
                    if hasattr(a, "ptr") and hasattr(a, "cp") and hasattr(a, "h"):
                        v = self._v4(a)
                        v[:] = v.astype(np.float16).astype(np.float32)
            return out
        inner.__name__, inner.__doc__, inner.__wrapped__ = fn.__name__, fn.__doc__, fn
        return inner

    for name, fn in list(vars(cls).items()):
        if callable(fn) and not name.startswith("_") and name not in _NOT_OPS:
            setattr(cls, name, wrap(fn))
    return cls


@_fp16_storage
class FakeRuntime:
    def __init__(self, enforce_device_limits=True, fp16=False):
        self.bufs, self.layers, self.launches = {}, [], 0
        self.enforce = enforce_device_limits
        self.fp16 = fp16        # simulate fp16 storage + fp16 tensor-core operands (fp32 accumulation): precision forecasts without a GPU
        self._next = 0x1000
        self._bases = []
        self._flag = False
        self._rec = None          # list of (method, args) while a capture is open
        self.graphs = []
        self.rescaled = 0         # absmax calls that reported a value above fp16

    # ---- graph capture: like a CUDA stream capture, calls are recorded with their argument VALUES and not executed
    def capture_begin(self):
        assert self._rec is None
        self._rec = []

    def capture_end(self):
        self.graphs.append(self._rec)
        self._rec = None
        return len(self.graphs) - 1

    def graph_launch(self, g):
        assert self._rec is None and self.graphs[g] is not None
        for name, args, kw in self.graphs[g]:
            getattr(self, name)(*args, **kw)

    def graph_destroy(self, g):
        self.graphs[g] = None

    def _recording(self, name, *args, **kw):
        if self._rec is None:
            return False
        self._rec.append((name, args, kw))
        return True

    def _store(self, y, values):
        if np.abs(values).max(initial=0.0) > 65504.0 or not np.isfinite(values).all():
            self._flag = True
        self._view(y)[...] = values

    def absmax(self, t):
        assert self._rec is None
        m = float(np.abs(self._view(t)).max())
        if m > 65504.0:
            self.rescaled += 1
            return float("inf")    # what the fp16 buffer of the device would report
        return m

    def overflow(self):
        assert self._rec is None
        f, self._flag = self._flag, False
        return f

    def close(self):
        self.bufs.clear()

    def _handle(self, nbytes):
        h = self._next
        self._next += (int(nbytes) + 0x1fff) // 0x1000 * 0x1000     # address space like the device: pointer arithmetic stays inside a buffer
        self._bases.append(h)
        return h

    def alloc(self, nbytes):
        assert self._rec is None, "allocation during graph capture"
        h = self._handle(nbytes)
        self.bufs[h] = np.zeros(int(nbytes) // 2, np.float32)   # one fp32 per fp16 element of the real buffer
        return h

    def free(self, ptr):
        assert self._rec is None, "free during graph capture"
        del self.bufs[ptr]
        self._bases.remove(ptr)

    def upload_f32(self, arr):
        assert self._rec is None, "upload during graph capture"
        arr = np.array(arr, np.float32)
        h = self._handle(arr.size * 4)
        self.bufs[h] = arr
        return h

    def _resolve(self, ptr):
        """device pointer -> (buffer, element offset); one element per 2 bytes"""
        import bisect

        base = self._bases[bisect.bisect_right(self._bases, ptr) - 1]
        off = ptr - base
        assert off % 2 == 0 and off // 2 <= self.bufs[base].size
        return self.bufs[base], off // 2

    def _v4(self, t):
        """[n, h, w, channels from the view's first channel on] (channel-slice views advance the pointer by 2*c0 bytes)."""
        arr, off = self._resolve(t.ptr)
        c0 = off % t.cp                     # tensors start at their buffer: offset = whole images (+ a channel offset)
        n = getattr(t, "n", 1)
        return arr[off - c0: off - c0 + n * t.h * t.w * t.cp].reshape(n, t.h, t.w, t.cp)[:, :, :, c0:]

    def _view(self, t):
        v = self._v4(t)
        return v[0] if v.shape[0] == 1 else v.reshape(v.shape[0] * v.shape[1], v.shape[2], v.shape[3])   # element-wise ops: images stacked on rows

    def conv_create(self, w, bias, cout, cin, cin_pitch, kh, kw, stride, pad_t, pad_l, dil, groups, transposed):
        w = np.array(w, np.float32)
        assert cin_pitch >= cin and cin_pitch % 8 == 0
        if self.enforce:   # the restrictions of vsr_rt_conv_create (engine.cu), so that violations show up on the CPU
            if transposed:
                assert (kh, kw, stride, groups) == (2, 2, 2, 1) and w.shape == (cin, cout, 2, 2)
            elif groups > 1:
                assert groups == cin == cout and kh == kw and dil == 1 and w.shape == (cin, 1, kh, kw)
                assert pad_t == pad_l
            elif cin < 16 or cout < 8:
                assert dil == 1 and w.shape == (cout, cin, kh, kw)
            elif stride == 2:
                assert (kh, kw, pad_t, pad_l, dil) == (3, 3, 1, 1, 1) and cin_pitch % 16 == 0 and cout % 8 == 0
            else:
                assert stride == 1 and cout % 8 == 0 and w.shape == (cout, cin, kh, kw)
                cin_k = (cin + 63) // 64 * 64          # pack_conv_general: whole 64-channel TMA boxes per tap, at most 81 taps
                assert kh * kw <= 81 and (cin_k <= cin_pitch or cin_pitch % 64 == 0), "a channel slice must end on the tensor's 64-channel grid"
        if self.fp16 and not transposed and groups == 1 and cin >= 16 and cout >= 8:      # tensor-core layers hold fp16 weights; the direct kernels fp32
            w = w.astype(np.float16).astype(np.float32)
        self.layers.append(dict(w=torch.from_numpy(w), b=torch.from_numpy(np.array(bias, np.float32)), cout=cout, cin=cin, kh=kh, kw=kw,
                                stride=stride, pad_t=pad_t, pad_l=pad_l, dil=dil, groups=groups, transposed=transposed, cin_pitch=cin_pitch))
        return len(self.layers) - 1

    def conv_create_split(self, w, bias, cout, cin, cin_pitch, kh, kw, pad_t, pad_l, dil):
        """hi + lo fp16 weights: in fp16 mode the layer sees fp16(w) + fp16(w - fp16(w)), otherwise w"""
        w = np.array(w, np.float32)
        assert cout >= 8 and cout % 8 == 0 and cin >= 16 and 2 * kh * kw <= 81
        if self.fp16:
            hi = w.astype(np.float16).astype(np.float32)
            w = hi + (w - hi).astype(np.float16).astype(np.float32)
        f, self.fp16 = self.fp16, False
        try:
            return self.conv_create(w, bias, cout, cin, cin_pitch, kh, kw, 1, pad_t, pad_l, dil, 1, False)
        finally:
            self.fp16 = f

    def conv_ex(self, lid, x, y, relu, out_coff=0, crop=None):
        """crop None: a plain conv of any kernel family (possibly into a channel slice); (top, left): the cropped-store path."""
        if not self._recording("conv_ex", lid, x, y, relu, out_coff, crop):
            self._conv(lid, x, y, relu, 1.0, 1.0, out_coff, crop)

    def conv(self, lid, x, y, relu, alpha=1.0, bias_scale=1.0):
        if not self._recording("conv", lid, x, y, relu, alpha, bias_scale):
            self._conv(lid, x, y, relu, alpha, bias_scale, 0, None)

    def _conv(self, lid, x, y, relu, alpha, bias_scale, out_coff, crop):
        L = self.layers[lid]
        if self.enforce:
            # the layer was packed for ONE input pitch (TMA strides / kernel arguments are built from it): a tensor of another pitch would be
            # read with the wrong strides on the device, silently; 16-byte stores need aligned bases, pitches and channel offsets
            assert x.cp == L["cin_pitch"], f"conv created for input pitch {L['cin_pitch']}, called with {x.cp}"
            assert x.ptr % 16 == 0 and y.ptr % 16 == 0 and y.cp % 8 == 0 and out_coff % 8 == 0, "16-byte accesses would fault on the device"
        if self.enforce and not (L["groups"] == 1 and not L["transposed"] and L["cin"] >= 16 and L["cout"] >= 8):
            # the run-time checks of vsr_rt_conv_ex for the depthwise / direct / transposed kernels: plain output tensors, symmetric padding
            assert out_coff == 0 and y.cp % 8 == 0 and crop is None, "direct / depthwise / transposed convs store whole tensors"
            if L["groups"] > 1:
                assert y.cp == x.cp
            else:
                assert y.cp >= _r8(L["cout"])
            if not L["transposed"]:
                assert (y.h, y.w) == ((x.h + 2 * L["pad_t"] - L["kh"]) // L["stride"] + 1, (x.w + 2 * L["pad_l"] - L["kw"]) // L["stride"] + 1), \
                    "these kernels derive the output size from symmetric padding"
        xin = torch.from_numpy(self._v4(x)[:, :, :, : L["cin"]].copy()).permute(0, 3, 1, 2)
        if crop is not None:   # 'same' conv on the (padded) input grid, keep the window [crop, crop + y.hw)
            assert not L["transposed"] and L["groups"] == 1 and L["cin"] >= 16 and L["cout"] >= 8, "cropped output needs a tensor-core conv"
            p = (L["kh"] - 1) * L["dil"] // 2
            assert (L["pad_t"], L["pad_l"]) == (p, p)
            if L["stride"] == 2:
                assert x.h % 2 == 0 and x.w % 2 == 0
            full = F.conv2d(F.pad(xin, (p, p, p, p)), L["w"], None, stride=L["stride"], dilation=L["dil"])
            out = full[:, :, crop[0]:crop[0] + y.h, crop[1]:crop[1] + y.w]
        elif L["transposed"]:
            out = F.conv_transpose2d(xin, L["w"], None, stride=2)
        else:
            eff_h, eff_w = (L["kh"] - 1) * L["dil"] + 1, (L["kw"] - 1) * L["dil"] + 1
            pad_b = max((y.h - 1) * L["stride"] + eff_h - x.h - L["pad_t"], 0)
            pad_r = max((y.w - 1) * L["stride"] + eff_w - x.w - L["pad_l"], 0)
            if self.enforce and L["stride"] == 2 and L["groups"] == 1 and L["cin"] >= 16:
                assert x.h % 2 == 0 and x.w % 2 == 0
            xin = F.pad(xin, (L["pad_l"], pad_r, L["pad_t"], pad_b))
            out = F.conv2d(xin, L["w"], None, stride=L["stride"], dilation=L["dil"], groups=L["groups"])
        assert out.shape[2:] == (y.h, y.w), (out.shape, y.h, y.w)
        out = out * alpha + (L["b"] * bias_scale)[None, :, None, None]
        if relu:
            out = out.relu()
        res = out.permute(0, 2, 3, 1).numpy()
        if np.abs(res).max(initial=0.0) > 65504.0 or not np.isfinite(res).all():
            self._flag = True
        v = self._v4(y)
        if out_coff == 0 and crop is None and _r8(L["cout"]) >= y.c:
            v[:] = 0           # the device kernels write zeros into the channel padding of a plain output tensor
        v[:, :, :, out_coff:out_coff + L["cout"]] = res
        self.launches += 1

    # ---- mobile detector
    def hswish_affine(self, x, y, inv_scale_in, a, c):
        if self._recording("hswish_affine", x, y, inv_scale_in, a, c):
            return
        v = self._view(x) * inv_scale_in
        self._store(y, a * (v * np.clip(v + 3.0, 0.0, 6.0) / 6.0) + c)
        self.launches += 1

    def se_create(self, w1, b1, w2, b2, slope, offset, residual):
        self.layers.append(dict(se=(np.array(w1, np.float32), np.array(b1, np.float32), np.array(w2, np.float32), np.array(b2, np.float32), slope, offset,
                                    residual)))
        return len(self.layers) - 1

    def se_gate(self, se_id, x, inv_scale, gate_ptr):
        if self._recording("se_gate", se_id, x, inv_scale, gate_ptr):
            return
        w1, b1, w2, b2, slope, offset, residual = self.layers[se_id]["se"]
        C = w1.shape[1]
        mean = self._view(x)[:, :, :C].reshape(-1, C).mean(0) * inv_scale
        g = np.clip(slope * (w2 @ np.maximum(w1 @ mean + b1, 0) + b2) + offset, 0.0, 1.0)
        out = np.zeros_like(self.bufs[gate_ptr])
        out[:C] = 1.0 + g if residual else g
        self.bufs[gate_ptr][:] = out
        self.launches += 2

    # ---- RAFT (csrc/pp_ops.cuh)
    def frames(self, frames_bgr, y):
        assert self._rec is None
        v = self._v4(y)
        v[:] = 0
        for t, f in enumerate(frames_bgr):
            v[t, :, :, :3] = (f[:, :, ::-1].astype(np.float32) / np.float32(255)) * 2 - 1
        self.launches += 1

    def instnorm(self, x, y, relu):
        assert self._rec is None
        v = self._v4(x)
        m = v.mean((1, 2), keepdims=True)
        r = (v - m) / np.sqrt(v.var((1, 2), keepdims=True) + 1e-5)
        self._v4(y)[:] = np.maximum(r, 0) if relu else r
        self.launches += 2

    def context_split(self, x, net, inp):
        assert self._rec is None
        v = self._v4(x)
        self._v4(net)[..., :128] = np.tanh(v[..., :128])
        self._v4(inp)[..., :128] = np.maximum(v[..., 128:256], 0)
        self.launches += 1

    def _raw(self, ptr, count):
        arr, off = self._resolve(ptr)
        return arr[off: off + count]

    def corr_volume(self, f1_ptr, f2_ptr, hh, ww, c, out_ptr, out_pitch):
        assert self._rec is None
        hw = hh * ww
        a = self._raw(f1_ptr, hw * c).reshape(hw, c)
        b = self._raw(f2_ptr, hw * c).reshape(hw, c)
        out = self._raw(out_ptr, hw * out_pitch).reshape(hw, out_pitch)
        out[:, :hw] = (a @ b.T) / np.sqrt(np.float32(c))
        if self.fp16:
            out[:] = out.astype(np.float16).astype(np.float32)
        self.launches += 1

    def corr_pool(self, in_ptr, rows, h2, w2, pitch_in, out_ptr, pitch_out):
        assert self._rec is None
        src = self._raw(in_ptr, rows * pitch_in).reshape(rows, pitch_in)[:, : h2 * w2].reshape(rows, h2, w2)
        oh, ow = h2 // 2, w2 // 2
        p = src[:, : 2 * oh, : 2 * ow].reshape(rows, oh, 2, ow, 2).mean((2, 4))
        dst = self._raw(out_ptr, rows * pitch_out).reshape(rows, pitch_out)
        dst[:, : oh * ow] = p.reshape(rows, -1)
        if self.fp16:
            dst[:] = dst.astype(np.float16).astype(np.float32)
        self.launches += 1

    def corr_lookup(self, levels, flow32, hh, ww, pixels, out):
        if self._recording("corr_lookup", levels, flow32, hh, ww, pixels, out):
            return
        flow = self._raw32(flow32, pixels * 2).reshape(pixels, 2)
        ys, xs = np.divmod(np.arange(pixels) % (hh * ww), ww)
        res = np.zeros((pixels, 324), np.float32)
        d = np.arange(-4, 5, dtype=np.float32)
        for l, (ptr, H, W, pitch) in enumerate(levels):
            m = self._raw(ptr, pixels * pitch).reshape(pixels, pitch)[:, : H * W].reshape(pixels, H, W)
            cx, cy = (xs + flow[:, 0]) / 2 ** l, (ys + flow[:, 1]) / 2 ** l
            X = cx[:, None, None] + d[None, :, None] + np.zeros((1, 1, 9), np.float32)     # (i, j) -> x + d[i]
            Y = cy[:, None, None] + d[None, None, :] + np.zeros((1, 9, 1), np.float32)     #           y + d[j]
            x0, y0 = np.floor(X).astype(np.int64), np.floor(Y).astype(np.int64)
            ax, ay = X - x0, Y - y0
            acc = np.zeros_like(X)
            rows = np.arange(pixels)[:, None, None]
            for dy_, wy in ((0, 1 - ay), (1, ay)):
                for dx_, wx in ((0, 1 - ax), (1, ax)):
                    yy, xx = y0 + dy_, x0 + dx_
                    ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
                    acc += np.where(ok, m[rows, np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], 0) * wy * wx
            res[:, l * 81:(l + 1) * 81] = acc.reshape(pixels, 81)
        v = self._v4(out).reshape(pixels, -1)
        v[:, :324] = res
        self.launches += 1

    def gru_rh(self, r, hsrc, out):
        if self._recording("gru_rh", r, hsrc, out):
            return
        self._v4(out)[..., :128] = self._v4(hsrc)[..., :128] / (1 + np.exp(-self._v4(r)[..., :128]))
        self.launches += 1

    def gru_update(self, z, q, hio):
        if self._recording("gru_update", z, q, hio):
            return
        zz = 1 / (1 + np.exp(-self._v4(z)[..., :128]))
        h = self._v4(hio)
        h[..., :128] = (1 - zz) * h[..., :128] + zz * np.tanh(self._v4(q)[..., :128])
        self.launches += 1

    def flow_update(self, flow32, delta, flow16, dst_a, dst_b, coff, add):
        if self._recording("flow_update", flow32, delta, flow16, dst_a, dst_b, coff, add):
            return
        n = flow16.pixels
        f = self._raw32(flow32, n * 2).reshape(n, 2)          # strided view into the buffer: in-place updates land in it
        if add:
            f += self._v4(delta).reshape(n, -1)[:, :2]
        v = self._v4(flow16).reshape(n, -1)
        v[:] = 0
        v[:, :2] = f
        for dst in (dst_a, dst_b):
            if dst is not None:
                self._v4(dst).reshape(n, -1)[:, coff:coff + 2] = f
        self.launches += 1

    def convex_upsample(self, flow32, mask, n, hh, ww, out32):
        assert self._rec is None
        f = torch.from_numpy(self._raw32(flow32, n * hh * ww * 2).reshape(n, hh, ww, 2).copy()).permute(0, 3, 1, 2)
        m = torch.from_numpy(self._v4(mask)[..., :576].copy()).permute(0, 3, 1, 2).reshape(n, 1, 9, 8, 8, hh, ww)
        m = torch.softmax(m, 2)
        up = F.unfold(8 * f, [3, 3], padding=1).view(n, 2, 9, 1, 1, hh, ww)
        out = torch.sum(m * up, 2).permute(0, 1, 4, 2, 5, 3).reshape(n, 2, 8 * hh, 8 * ww)
        self._raw32(out32, out.numel())[:] = out.numpy().reshape(-1)
        self.launches += 1

    def download_f32(self, ptr, shape):
        return self._raw32(ptr, int(np.prod(shape))).reshape(shape).copy()

    def zero(self, ptr, nbytes):
        self.bufs[ptr][: nbytes // 2] = 0

    # ---- ProPainter flow completion (P4): transcriptions of the pp_ops.cuh kernels
    def rfc_input(self, flow32, mask_u8, n, hh, ww, reverse, out):
        assert self._rec is None
        f = self._raw32(flow32, n * 2 * hh * ww).reshape(n, 2, hh, ww)
        if reverse:
            f = f[::-1]
        m = (self.bufs[mask_u8][: hh * ww].reshape(hh, ww) > 0).astype(np.float32)
        v = self._v4(out)
        v[:] = 0
        v[..., 0], v[..., 1], v[..., 2] = f[:, 0] * (1 - m), f[:, 1] * (1 - m), m
        self.launches += 1

    def pad_replicate(self, x, y, top, left):
        assert self._rec is None
        self._v4(y)[:] = np.pad(self._v4(x), ((0, 0), (top, y.h - x.h - top), (left, y.w - x.w - left), (0, 0)), mode="edge")
        self.launches += 1

    def leaky(self, x, slope):
        assert self._rec is None
        v = self._v4(x)
        v[:] = np.where(v > 0, v, v * slope)
        self.launches += 1

    def temporal_taps(self, x, y):
        assert self._rec is None
        v, o = self._v4(x), self._v4(y)
        o[:] = 0
        T = v.shape[0]
        for k in range(3):
            for t in range(T):
                ts = t + 2 * (k - 1)
                if 0 <= ts < T:
                    o[t, :, :, k * x.cp: (k + 1) * x.cp] = v[ts]
        self.launches += 1

    def deform_cols(self, xa, ca, xb, c, groups, om, max_residue, flow32, cols):
        assert self._rec is None
        H, W = xa.h, xa.w
        a = self._v4(xa)[..., :ca]
        x = a if xb is None else np.concatenate([a, self._v4(xb)[..., : c - ca]], -1)       # [n,H,W,C]
        o = self._v4(om)
        n = x.shape[0]
        gk = groups * 9
        off = max_residue * np.tanh(o[..., : 2 * gk]).reshape(n, H, W, gk, 2)
        if flow32:
            fl = self._flow_at(flow32, n * H, W).reshape(n, H, W, 1, 2)
            off = off + fl[..., ::-1]                                                            # (dy, dx) += (flow.y, flow.x)
        msk = 1 / (1 + np.exp(-o[..., 2 * gk: 3 * gk]))
        ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
        out = np.zeros((n, H, W, 9, c), np.float32)
        cpg = c // groups
        for g in range(groups):
            xg = x[..., g * cpg:(g + 1) * cpg]
            for k in range(9):
                ky, kx = divmod(k, 3)
                sy = ys[None] - 1 + ky + off[:, :, :, g * 9 + k, 0]
                sx = xs[None] - 1 + kx + off[:, :, :, g * 9 + k, 1]
                y0, x0 = np.floor(sy), np.floor(sx)
                ay, ax = sy - y0, sx - x0
                y0, x0 = y0.astype(np.int64), x0.astype(np.int64)
                acc = np.zeros((n, H, W, cpg), np.float32)
                nn = np.arange(n)[:, None, None]
                for dy, wy in ((0, 1 - ay), (1, ay)):
                    for dx, wx in ((0, 1 - ax), (1, ax)):
                        yy, xx = y0 + dy, x0 + dx
                        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
                        acc += np.where(ok[..., None], xg[nn, np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], 0) * (wy * wx)[..., None]
                out[:, :, :, k, g * cpg:(g + 1) * cpg] = acc * msk[:, :, :, g * 9 + k, None]
        self._v4(cols)[..., : 9 * c] = out.reshape(n, H, W, 9 * c)
        self.launches += 1

    def rfc_combine(self, pred, flow32, mask_u8, n, hh, ww, reverse, out32):
        assert self._rec is None
        p = self._v4(pred)[..., :2]
        if reverse:
            p = p[::-1]
        f = self._raw32(flow32, n * 2 * hh * ww).reshape(n, 2, hh, ww)
        m = (self.bufs[mask_u8][: hh * ww].reshape(hh, ww) > 0).astype(np.float32)
        res = p.transpose(0, 3, 1, 2) * m + f * (1 - m)
        self._raw32(out32, res.size)[:] = res.reshape(-1)
        self.launches += 1

    def upsample2x(self, x, y):
        assert self._rec is None
        v = torch.from_numpy(self._v4(x).copy()).permute(0, 3, 1, 2)
        self._v4(y)[:] = F.interpolate(v, scale_factor=2, mode="bilinear", align_corners=True).permute(0, 2, 3, 1).numpy()
        self.launches += 1

    # ---- ProPainter generator, front half (P6): transcriptions of the pp_ops.cuh kernels
    def gen_input(self, state, mask_u8, ids_dev, n, out):
        assert self._rec is None
        ids = self.bufs[ids_dev][:n].astype(np.int64)
        st = self._v4(state)
        m = (self.bufs[mask_u8][: state.h * state.w].reshape(state.h, state.w) > 0).astype(np.float32)
        o = self._v4(out)
        o[:] = 0
        o[..., :3] = st[ids][..., :3]
        o[..., 3] = m
        o[..., 4] = st[ids][..., 3]
        self.launches += 1

    def flow_down4(self, flow32, ids_dev, n, hh, ww, out32):
        assert self._rec is None
        ids = self.bufs[ids_dev][:n].astype(np.int64)
        total = int(ids.max()) + 1
        f = self._raw32(flow32, total * 2 * hh * ww).reshape(total, 2, hh, ww)[ids]
        blk = (f[:, :, 1::4, 1::4] + f[:, :, 1::4, 2::4] + f[:, :, 2::4, 1::4] + f[:, :, 2::4, 2::4]) * 0.25 * 0.25
        res = blk.transpose(0, 2, 3, 1).reshape(-1)
        self._raw32(out32, res.size)[:] = res
        self.launches += 1

    def prop_masks(self, gen_in, out):
        assert self._rec is None
        g = self._v4(gen_in)
        o = self._v4(out)
        o[:] = 0
        o[..., 0], o[..., 1] = g[:, ::4, ::4, 3], g[:, ::4, ::4, 4]
        self.launches += 1

    def _flow_at(self, ptr, hh, ww):
        """fp32 [P][2] flow at a byte-offset pointer (one float per two slots, like every fp32 buffer of this stand-in)"""
        return self._raw32(ptr, hh * ww * 2).reshape(hh, ww, 2)

    def featprop_cond(self, prop, cur, flow_prop, flow_check, masks, cond):
        assert self._rec is None
        H, W, C = cur.h, cur.w, cur.cp
        fp, fc = self._flow_at(flow_prop, H, W), self._flow_at(flow_check, H, W)
        pv, cv, mk = self._v4(prop)[0], self._v4(cur)[0], self._v4(masks)[0]
        ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
        sx, sy = xs + fp[..., 0], ys + fp[..., 1]
        x0, y0 = np.floor(sx), np.floor(sy)
        ax, ay = sx - x0, sy - y0
        x0, y0 = x0.astype(np.int64), y0.astype(np.int64)
        warped = np.zeros((H, W, C), np.float32)
        b = np.zeros((H, W, 2), np.float32)
        for dy, wy in ((0, 1 - ay), (1, ay)):
            for dx, wx in ((0, 1 - ax), (1, ax)):
                yy, xx = y0 + dy, x0 + dx
                ok = ((yy >= 0) & (yy < H) & (xx >= 0) & (xx < W))[..., None]
                wgt = (wy * wx)[..., None]
                warped += np.where(ok, pv[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], 0) * wgt
                b += np.where(ok, fc[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], 0) * wgt
        d = fp + b
        valid = (d ** 2).sum(-1) < 0.01 * ((fp ** 2).sum(-1) + (b ** 2).sum(-1)) + 0.5
        o = self._v4(cond)[0]
        o[..., :C] = cv
        o[..., C:2 * C] = warped
        o[..., 2 * C:2 * C + 2] = fp
        o[..., 2 * C + 2] = valid
        o[..., 2 * C + 3:2 * C + 5] = mk[..., :2]
        self.launches += 1

    def write_extra(self, src, dst, coff, nch):
        assert self._rec is None
        self._v4(dst)[..., coff:coff + nch] = self._v4(src)[..., :nch]
        self.launches += 1

    # ---- ProPainter generator, back half: transcriptions of the pp_ops.cuh kernels
    def upload_ints(self, arr):
        return self.upload_bytes(np.asarray(arr, np.int32))

    def unfold7s3(self, x, out, gelu=False):
        assert self._rec is None
        from math import erf

        v = self._v4(x)[..., : x.cp]
        n, h, w, C = v.shape
        fh, fw = (h + 6 - 7) // 3 + 1, (w + 6 - 7) // 3 + 1
        p = np.pad(v, ((0, 0), (3, 3 + 3), (3, 3 + 3), (0, 0)))
        o = self._v4(out)
        for k in range(49):
            ky, kx = divmod(k, 7)
            blk = p[:, ky: ky + 3 * fh: 3, kx: kx + 3 * fw: 3]
            if gelu:
                blk = 0.5 * blk * (1 + np.vectorize(erf)(blk * 0.70710678))
            o[..., k * C:(k + 1) * C] = blk
        self.launches += 1

    def fold7s3(self, tok, out, channels, normalise):
        assert self._rec is None
        t = self._v4(tok)
        n, fh, fw, _ = t.shape
        h, w, C = out.h, out.w, channels
        acc = np.zeros((n, h + 12, w + 12, C), np.float32)
        cnt = np.zeros((h + 12, w + 12), np.float32)
        for k in range(49):
            ky, kx = divmod(k, 7)
            acc[:, ky: ky + 3 * fh: 3, kx: kx + 3 * fw: 3] += t[..., k * C:(k + 1) * C]
            cnt[ky: ky + 3 * fh: 3, kx: kx + 3 * fw: 3] += 1
        acc, cnt = acc[:, 3: 3 + h, 3: 3 + w], cnt[3: 3 + h, 3: 3 + w]
        if normalise:
            acc = acc / np.maximum(cnt, 1)[None, :, :, None]
        self._v4(out)[..., :C] = acc
        self.launches += 1

    def layernorm(self, x, gamma, beta, out):
        assert self._rec is None
        v = self._v4(x)
        m = v.mean(-1, keepdims=True)
        r = 1 / np.sqrt(np.maximum((v * v).mean(-1, keepdims=True) - m * m, 0) + 1e-5)
        self._v4(out)[:] = (v - m) * r * self.bufs[gamma][: x.cp] + self.bufs[beta][: x.cp]
        self.launches += 1

    def pool4(self, x, w_dev, b_dev, out):
        assert self._rec is None
        v = self._v4(x)
        n, H, W, C = v.shape
        ph, pw = H // 4, W // 4
        wt = self.bufs[w_dev][: C * 16].reshape(C, 4, 4)
        blk = v[:, : 4 * ph, : 4 * pw].reshape(n, ph, 4, pw, 4, C)
        self._v4(out)[:] = np.einsum("nyaxbc,cab->nyxc", blk, wt) + self.bufs[b_dev][:C]
        self.launches += 1

    def window_attention(self, q, k, v, kp, vp, valid_dev, n_valid, tind_dev, n_tind, masked_dev, out):
        assert self._rec is None
        Q, K, V, KP, VP = (self._v4(t) for t in (q, k, v, kp, vp))
        T, Hn, Wn, C = Q.shape
        valid = self.bufs[valid_dev][:n_valid].astype(np.int64)
        tind = self.bufs[tind_dev][:n_tind].astype(np.int64)
        nwh, nww = Hn // 5, Wn // 9
        masked = self.bufs[masked_dev][: nwh * nww] > 0
        eh, ew = 3, 5
        O = self._v4(out)
        O[:] = 0
        own = [(s // 9, s % 9) for s in range(45)]
        rolled = []
        for idv in valid:
            rr, o = divmod(int(idv), 45)
            rolled.append((o // 9, o % 9, -eh if rr < 2 else eh, ew if rr & 1 else -ew))
        for win in range(nwh * nww):
            wy0, wx0 = (win // nww) * 5, (win % nww) * 9
            ys = np.array([wy0 + a for a, _ in own])
            xs = np.array([wx0 + b for _, b in own])
            ry = np.array([(wy0 + a - sy) % Hn for a, _, sy, _ in rolled])
            rx = np.array([(wx0 + b - sx) % Wn for _, b, _, sx in rolled])
            for hd in range(C // 128):
                cs = slice(hd * 128, (hd + 1) * 128)
                for tq in range(T):
                    qq = Q[tq, ys, xs, cs]                                     # [45, 128]
                    if masked[win]:
                        kk = np.concatenate([np.concatenate([K[t, ys, xs, cs], K[t, ry, rx, cs], KP[t].reshape(-1, C)[:, cs]]) for t in tind])
                        vv = np.concatenate([np.concatenate([V[t, ys, xs, cs], V[t, ry, rx, cs], VP[t].reshape(-1, C)[:, cs]]) for t in tind])
                    else:
                        kk, vv = K[tq, ys, xs, cs], V[tq, ys, xs, cs]
                    a = (qq @ kk.T) * 0.08838834764831845
                    a = np.exp(a - a.max(1, keepdims=True))
                    O[tq, ys, xs, cs] = (a / a.sum(1, keepdims=True)) @ vv
        self.launches += 1

    def pred_to_rgb8(self, x):
        assert self._rec is None
        v = self._v4(x)[..., :3]
        return np.clip((np.tanh(v) + 1) / 2 * 255, 0, 255).astype(np.uint8)

    # ---- ProPainter image propagation (P5)
    def upload_bytes(self, arr):
        arr = np.ascontiguousarray(arr)
        h = self._handle(max(arr.nbytes, 16))
        if arr.dtype == np.float32:        # one slot per 2 bytes: an fp32 element occupies two slots (value, unused), so byte offsets stay valid
            slots = np.zeros(2 * arr.size, np.float32)
            slots[0::2] = arr.reshape(-1)
            self.bufs[h] = slots
        else:                              # u8 masks: no pointer arithmetic on them, one slot per element
            self.bufs[h] = arr.astype(np.float32).reshape(-1)
        return h

    def copy_bytes(self, src, dst, nbytes):
        assert self._rec is None and src % 2 == 0 and dst % 2 == 0 and nbytes % 2 == 0
        a, ao = self._resolve(src)
        b, bo = self._resolve(dst)
        b[bo: bo + nbytes // 2] = a[ao: ao + nbytes // 2]          # slots are 2 bytes of device memory each, whatever they hold

    def upload_to(self, ptr, arr):
        arr = np.ascontiguousarray(arr)
        need = 2 * arr.size if arr.dtype == np.float32 else arr.size
        if self.bufs[ptr].size < need:            # alloc() sized the slots for fp16 data; byte / int payloads need one slot per element here
            self.bufs[ptr] = np.zeros(need, np.float32)
        buf = self.bufs[ptr]
        if arr.dtype == np.float32:
            buf[: 2 * arr.size: 2] = arr.reshape(-1)
        else:
            buf[: arr.size] = arr.reshape(-1)

    def _raw32(self, ptr, count):
        arr, off = self._resolve(ptr)
        return arr[off: off + 2 * count: 2]

    def prop_state(self, frames, mask_u8, prop, out):
        assert self._rec is None
        f = self._v4(frames)
        m = (self.bufs[mask_u8][: frames.h * frames.w].reshape(1, frames.h, frames.w) > 0)
        o = self._v4(out)
        o[:] = f
        if prop is None:
            o[..., :3] = np.where(m[..., None], 0.0, f[..., :3])
            o[..., 3] = m
        else:
            p = self._v4(prop)
            o[..., :3] = np.where(m[..., None], p[..., :3], f[..., :3])
            o[..., 3] = p[..., 3]
        self.launches += 1

    def img_prop_step(self, prev, cur, flow_prop, flow_check, out):
        """numpy transcription of pp_img_prop_step_kernel (NOT of the torch reference): hand-written bilinear / nearest sampling, so
        that the CPU test checks the kernel's arithmetic against the oracle."""
        assert self._rec is None
        H, W = cur.h, cur.w
        fp = self._raw32(flow_prop, 2 * H * W).reshape(2, H, W)
        fc = self._raw32(flow_check, 2 * H * W).reshape(2, H, W)
        pv, cu = self._v4(prev)[0], self._v4(cur)[0]
        ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
        sx, sy = xs + fp[0], ys + fp[1]

        def bilinear(m):
            x0, y0 = np.floor(sx), np.floor(sy)
            ax, ay = sx - x0, sy - y0
            x0, y0 = x0.astype(np.int64), y0.astype(np.int64)
            acc = np.zeros((H, W), np.float32)
            for dy, wy in ((0, 1 - ay), (1, ay)):
                for dx, wx in ((0, 1 - ax), (1, ax)):
                    yy, xx = y0 + dy, x0 + dx
                    ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
                    acc += np.where(ok, m[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], 0) * wy * wx
            return acc

        bx, by = bilinear(fc[0]), bilinear(fc[1])
        dx, dy = fp[0] + bx, fp[1] + by
        valid = dx * dx + dy * dy < 0.01 * (fp[0] ** 2 + fp[1] ** 2 + bx * bx + by * by) + 0.5
        hole_there = bilinear(pv[..., 3]) > 0.1
        cur_hole = cu[..., 3] > 0.1
        fill = valid & ~hole_there
        nx, ny = np.rint(sx).astype(np.int64), np.rint(sy).astype(np.int64)       # round half to even, like nearbyintf
        inside = (nx >= 0) & (nx < W) & (ny >= 0) & (ny < H)
        warped = np.where(inside[..., None], pv[np.clip(ny, 0, H - 1), np.clip(nx, 0, W - 1), :3], 0)
        o = self._v4(out)[0]
        o[:] = cu
        take = cur_hole & fill
        o[..., :3] = np.where(take[..., None], warped, cu[..., :3])
        o[..., 3] = (cur_hole & ~fill).astype(np.float32)
        self.launches += 1

    # ---- LAMA-only entry points
    def pad(self, x, y, top, left, reflect=1):
        if self._recording("pad", x, y, top, left, reflect):
            return
        bottom, right = y.h - x.h - top, y.w - x.w - left
        if bottom < 0 or right < 0:                       # zero-mode window smaller than the input: a crop (out[oy][ox] = in[oy - top][ox - left])
            assert not reflect and top == 0 and left == 0
            self._v4(y)[:] = self._v4(x)[:, : y.h, : y.w]
        else:
            self._v4(y)[:] = np.pad(self._v4(x), ((0, 0), (top, bottom), (left, right), (0, 0)), mode="reflect" if reflect else "constant")
        self.launches += 1

    def zero_upsample(self, x, y):
        if self._recording("zero_upsample", x, y):
            return
        v = self._v4(y)
        v[:] = 0
        v[:, ::2, ::2] = self._v4(x)
        self.launches += 1

    def add_slices(self, relu, a, b, y, channels):
        if self._recording("add_slices", relu, a, b, y, channels):
            return
        r = self._view(a)[:, :, :channels] + self._view(b)[:, :, :channels]
        self._view(y)[:, :, :channels] = np.maximum(r, 0) if relu else r
        self.launches += 1

    def residual_add(self, x32, y, x, init):
        if self._recording("residual_add", x32, y, x, init):
            return
        master = self.bufs[x32]          # alloc() gave one fp32 per 2 bytes: the first pixels*cp entries are the fp32 master
        n = x.pixels * x.cp
        if init:
            master[:n] = self.bufs[x.ptr][:n]
        master[:n] += self.bufs[y.ptr][:n]
        self.bufs[x.ptr][:n] = master[:n]
        self.launches += 1

    def fft_r2c(self, x, y):
        if self._recording("fft_r2c", x, y):
            return
        f = np.fft.rfft2(self._v4(x)[..., : x.c].astype(np.float64), axes=(1, 2), norm="ortho")
        out = np.stack([f.real, f.imag], -1).reshape(f.shape[0], y.h, y.w, 2 * x.c)       # channel 2c + {re, im}
        self._v4(y)[..., : 2 * x.c] = out.astype(np.float32)
        self.launches += 3

    def fft_c2r(self, x, y):
        if self._recording("fft_c2r", x, y):
            return
        v = self._v4(x)[..., : 2 * y.c].astype(np.float64).reshape(-1, x.h, x.w, y.c, 2)
        out = np.fft.irfft2(v[..., 0] + 1j * v[..., 1], s=(y.h, y.w), axes=(1, 2), norm="ortho")
        self._v4(y)[..., : y.c] = out.astype(np.float32)
        self.launches += 3

    def lama_input(self, img, mask, y, slot=0):
        assert self._rec is None
        h, w = mask.shape
        m = (np.pad(mask, ((0, y.h - h), (0, y.w - w)), mode="symmetric") > 0).astype(np.float32)
        im = np.pad(img.astype(np.float32) / np.float32(255), ((0, y.h - h), (0, y.w - w), (0, 0)), mode="symmetric")
        v = self._v4(y)[slot]
        v[:] = 0
        v[:, :, :3] = im * (1 - m)[:, :, None]
        v[:, :, 3] = m
        self._staged = getattr(self, "_staged", {})
        self._staged[slot] = (img.copy(), mask.copy())
        self.launches += 1

    def lama_output(self, pred, ih, iw, slot=0):
        assert self._rec is None
        img, mask = self._staged[slot]
        m = (mask > 0).astype(np.float32)[:, :, None]
        res = m * self._v4(pred)[slot, :ih, :iw, :3] + (1 - m) * (img.astype(np.float32) / np.float32(255))
        self.launches += 1
        return np.clip(res * 255, 0, 255).astype(np.uint8)

    def elementwise(self, op, a, b, y, scale=0, shift=0, alpha=1.0, beta=1.0):
        if self._recording("elementwise", op, a, b, y, scale, shift, alpha, beta):
            return
        va = self._view(a)
        vb = self._view(b) if b is not None else None
        if op == 0:
            r = va * alpha + vb * beta
        elif op == 1:
            r = np.maximum(va, 0)
        elif op == 2:
            r = np.maximum(va * alpha + vb * beta, 0)
        elif op == 3:
            with np.errstate(over="ignore"):
                r = 1.0 / (1.0 + np.exp(-va * alpha))
        elif op in (4, 5):
            r = va * self.bufs[scale][None, None, : a.cp] + self.bufs[shift][None, None, : a.cp]
            if op == 5:
                r = np.maximum(r, 0)
        elif op == 6:
            r = va * alpha + beta
        elif op == 7:
            r = (va + vb) * alpha
        else:
            raise AssertionError(op)
        self._store(y, r)
        self.launches += 1

    def upsample(self, x, y, s):
        if self._recording("upsample", x, y, s):
            return
        self._v4(y)[:] = self._v4(x).repeat(s, axis=1).repeat(s, axis=2)
        self.launches += 1

    def maxpool(self, x, y):   # 2x2 stride 1, SAME: pad 0 top/left, 1 bottom/right (per image)
        if self._recording("maxpool", x, y):
            return
        v = self._v4(x)
        p = np.pad(v, ((0, 0), (0, 1), (0, 1), (0, 0)), constant_values=-np.inf)
        self._v4(y)[:] = np.maximum(np.maximum(p[:, :-1, :-1], p[:, 1:, :-1]), np.maximum(p[:, :-1, 1:], p[:, 1:, 1:]))
        self.launches += 1

    def copy_channels(self, src, dst, dst_off, channels):
        if self._recording("copy_channels", src, dst, dst_off, channels):
            return
        assert channels % 8 == 0 and dst_off % 8 == 0 and dst_off + channels <= dst.cp and channels <= src.cp
        self._view(dst)[:, :, dst_off:dst_off + channels] = self._view(src)[:, :, :channels]
        self.launches += 1

    def preprocess(self, img, inp, rh, rw, slot=0):
        assert self._rec is None
        x = D.preprocess(img)[0].permute(1, 2, 0).numpy()
        assert x.shape[:2] == (rh, rw)
        v = self._v4(inp)[slot]
        v[:] = 0
        v[:, :, :3] = x
        self.launches += 1

    def sync(self):
        pass

    def download_channel(self, t, ch, slot=0):
        return (self._v4(t)[slot, :, :, ch] / t.scale).astype(np.float32)

    def download(self, t):
        return self._view(t).copy()

    @property
    def launch_count(self):
        return self.launches
