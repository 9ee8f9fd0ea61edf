"""RAFT optical flow for the ProPainter path on the device-tensor runtime.

STATUS — read this first: the graph below is checked on the CPU against the oracle (oracle/raft_oracle.py, pinned to the
reference's flows) through the fp32 stand-in of the runtime (tests/test_raft_cpu.py), and its CUDA kernels
(csrc/pp_ops.cuh) compile for sm_100a, but round 1's GPU budget was spent before it existed: it has NOT run on a B200 yet and
nothing else in the package uses it (there is no PropainterInpaint class).  tests/test_gpu_raft.py is gated on
VSR_RUN_UNVALIDATED=1 for that reason.

What it mirrors: `RAFT_bi.forward` (backend/inpaint/video/model/modules/flow_comp_raft.py:39-55) -> `RAFT.forward`
(backend/inpaint/video/raft/raft.py:87-146, large model, 20 iterations, test mode) for every consecutive frame pair in both
directions.  Layout decisions:
* both encoders run once over all T frames (the forward pairs use frames 0..T-2 as image1, the backward pairs 1..T-1);
  batch-norm (cnet) is folded into the convs, instance norm (fnet) is a statistics + apply kernel pair per layer;
* the all-pairs correlation is one tensor-core GEMM per pair: a 1x1 "conv" of fmap1 whose weight matrix is fmap2 itself
  (NHWC features are already the K-major [Cout][K] layout), scaled by 1/sqrt(256) in the epilogue; 3 pooling launches
  build the pyramid; the 4 x 9 x 9 lookup is a gather kernel that keeps the reference's (dy -> x, dx -> y) offset order;
* the update block's state lives in two 384-channel tensors [h | inp | motion] and [r*h | inp | motion], so every GRU conv
  reads one tensor and `torch.cat` never materialises; the flow keeps an fp32 master next to its fp16 copies;
* one iteration (39 launches) is recorded as a CUDA graph and replayed; the mask head and the convex 8x up-sampling run
  once after the last iteration.
"""
import ctypes as C
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import _capi
from .dbnet import _Tensor, _r
from .lama_inpaint import _LamaRuntime, _view

ITERS = 20   # propainter_inpaint.py:156


def load_raft_weights(path_or_dict) -> Dict[str, np.ndarray]:
    """raft-things.pth (DataParallel-prefixed state dict; torch is only the container) or a name -> array dict."""
    if isinstance(path_or_dict, dict):
        return {k: np.asarray(v, np.float32) for k, v in path_or_dict.items()}
    import torch

    sd = torch.load(str(path_or_dict), map_location="cpu")
    return {(k[len("module."):] if k.startswith("module.") else k): v.float().numpy() for k, v in sd.items()}


class _RaftRuntime(_LamaRuntime):
    def conv_ex(self, lid, x, y, relu, out_coff=0, crop=None):
        """`crop=None`: a plain conv (any kernel family); a (top, left) pair selects the cropped-store tensor-core path."""
        c = crop if crop is not None else (0, 0)
        oh, ow = (y.h, y.w) if crop is not None else (0, 0)
        _capi.check(self.L.vsr_rt_conv_ex(self.h, lid, x.ptr, x.n, x.h, x.w, y.ptr, y.cp, out_coff, relu, 1.0, 1.0, c[0], c[1], oh, ow))

    def frames(self, frames_bgr: Sequence[np.ndarray], y):
        ptrs = (C.c_void_p * len(frames_bgr))(*[f.ctypes.data for f in frames_bgr])
        _capi.check(self.L.vsr_rt_pp_frames(self.h, ptrs, len(frames_bgr), y.h, y.w, y.ptr))

    def instnorm(self, x, y, relu):
        _capi.check(self.L.vsr_rt_instnorm(self.h, x.ptr, x.n, x.h * x.w, x.cp, relu, y.ptr))

    def context_split(self, x, net, inp):
        _capi.check(self.L.vsr_rt_context_split(self.h, x.ptr, x.pixels, net.ptr, net.cp, inp.ptr, inp.cp))

    def corr_volume(self, f1_ptr, f2_ptr, hh, ww, c, out_ptr, out_pitch):
        _capi.check(self.L.vsr_rt_corr_volume(self.h, f1_ptr, f2_ptr, hh, ww, c, out_ptr, out_pitch))

    def corr_pool(self, in_ptr, rows, h2, w2, pitch_in, out_ptr, pitch_out):
        _capi.check(self.L.vsr_rt_corr_pool(self.h, in_ptr, rows, h2, w2, pitch_in, out_ptr, pitch_out))

    def corr_lookup(self, levels, flow32, hh, ww, pixels, out):
        ptr = (C.c_uint64 * 4)(*[lv[0] for lv in levels])
        hs, ws, ps = ((C.c_int32 * 4)(*[lv[i] for lv in levels]) for i in (1, 2, 3))
        _capi.check(self.L.vsr_rt_corr_lookup(self.h, ptr, hs, ws, ps, flow32, hh, ww, pixels, out.ptr, out.cp))

    def gru_rh(self, r, hsrc, out):
        _capi.check(self.L.vsr_rt_gru_rh(self.h, r.ptr, r.cp, hsrc.ptr, hsrc.cp, out.ptr, out.cp, r.pixels))

    def gru_update(self, z, q, hio):
        _capi.check(self.L.vsr_rt_gru_update(self.h, z.ptr, z.cp, q.ptr, q.cp, hio.ptr, hio.cp, z.pixels))

    def flow_update(self, flow32, delta, flow16, dst_a, dst_b, coff, add):
        _capi.check(self.L.vsr_rt_flow_update(self.h, flow32, delta.ptr if delta is not None else 0, delta.cp if delta is not None else 0, flow16.ptr,
                                              dst_a.ptr if dst_a is not None else 0, dst_b.ptr if dst_b is not None else 0,
                                              dst_a.cp if dst_a is not None else 0, coff, flow16.pixels, 1 if add else 0))

    def convex_upsample(self, flow32, mask, n, hh, ww, out32):
        _capi.check(self.L.vsr_rt_convex_upsample(self.h, flow32, mask.ptr, mask.cp, n, hh, ww, out32))

    def download_f32(self, ptr, shape) -> np.ndarray:
        out = np.empty(shape, np.float32)
        _capi.check(self.L.vsr_rt_download(self.h, ptr, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def zero(self, ptr, nbytes):
        z = np.zeros(nbytes, np.uint8)
        _capi.check(self.L.vsr_rt_upload(self.h, ptr, z.ctypes.data_as(C.c_void_p), nbytes))


def _fold_bn(w, conv, norm):
    """conv (+bias) followed by eval-mode BatchNorm2d -> (weight, bias)."""
    s = w[f"{norm}.weight"] / np.sqrt(w[f"{norm}.running_var"] + np.float32(1e-5))
    return w[f"{conv}.weight"] * s[:, None, None, None], (w[f"{conv}.bias"] - w[f"{norm}.running_mean"]) * s + w[f"{norm}.bias"]


def _center3(wt):
    """1x1 stride-2 conv as the centre tap of a 3x3 stride-2 pad-1 conv (the tensor-core stride-2 path is 3x3 only)."""
    out = np.zeros(wt.shape[:2] + (3, 3), np.float32)
    out[:, :, 1, 1] = wt[:, :, 0, 0]
    return out


class _Prog:
    pass


class RaftFlow:
    """flows = RaftFlow(weights, device)(frames): bidirectional flows of consecutive BGR uint8 frames, like RAFT_bi."""

    def __init__(self, weights, device="cuda:0", runtime=None):
        self.w = load_raft_weights(weights)
        self._rt = runtime if runtime is not None else _RaftRuntime(device)
        self._layers: Dict[tuple, int] = {}
        self._progs: Dict[tuple, _Prog] = {}

    def __del__(self):
        rt = getattr(self, "_rt", None)
        if rt is not None:
            try:
                rt.close()
            except Exception:
                pass
            self._rt = None

    def _conv(self, key, weight, bias, cin_pitch, stride=1, pad=(0, 0)):
        k = (key, cin_pitch)
        if k not in self._layers:
            cout, cin, kh, kw = weight.shape
            if cout >= 8 and cout % 8:                      # 126 -> 128: the tensor-core conv stores 8-channel groups
                padn = _r(cout, 8) - cout
                weight = np.concatenate([weight, np.zeros((padn,) + weight.shape[1:], np.float32)])
                bias = np.concatenate([bias, np.zeros(padn, np.float32)])
            self._layers[k] = self._rt.conv_create(weight, bias, weight.shape[0], cin, cin_pitch, kh, kw, stride, pad[0], pad[1], 1, 1, False)
        return self._layers[k]

    # ------------------------------------------------------------------------------------------------ encoders
    def _encoder(self, prog, new, x, p, kind):
        """BasicEncoder.forward (extractor.py:118-190) on [T,H,W,8] -> [T,H/8,W/8,256]; kind 'instance' (fnet) or 'batch' (cnet)."""
        rt, w, run = self._rt, self.w, prog.eager.append

        def conv_norm(x, conv, norm, stride, relu, wt=None, ksize=None):
            wt = w[f"{conv}.weight"] if wt is None else wt
            b = w[f"{conv}.bias"]
            if kind == "batch" and norm is not None:
                s = w[f"{norm}.weight"] / np.sqrt(w[f"{norm}.running_var"] + np.float32(1e-5))
                wt, b = wt * s[:, None, None, None], (b - w[f"{norm}.running_mean"]) * s + w[f"{norm}.bias"]
            k = wt.shape[2]
            y = new(wt.shape[0], x.h // stride, x.w // stride)
            lid = self._conv((p, conv, kind), wt, b, x.cp, stride, (k // 2, k // 2))
            fused_relu = 1 if (relu and (kind == "batch" or norm is None)) else 0
            run(lambda: rt.conv_ex(lid, x, y, fused_relu))
            if kind == "instance" and norm is not None:
                z = new(y.c, y.h, y.w)
                run(lambda: rt.instnorm(y, z, 1 if relu else 0))
                return z
            return y

        def block(x, q, stride):
            y = conv_norm(x, f"{q}.conv1", f"{q}.norm1", stride, True)
            y = conv_norm(y, f"{q}.conv2", f"{q}.norm2", 1, True)
            if stride != 1:
                x = conv_norm(x, f"{q}.downsample.0", f"{q}.downsample.1", stride, False, wt=_center3(w[f"{q}.downsample.0.weight"]))
            out = new(y.c, y.h, y.w)
            run(lambda: rt.elementwise(2, x, y, out))
            return out

        x = conv_norm(x, f"{p}.conv1", f"{p}.norm1", 2, True)
        for layer, stride in (("layer1", 1), ("layer2", 2), ("layer3", 2)):
            x = block(x, f"{p}.{layer}.0", stride)
            x = block(x, f"{p}.{layer}.1", 1)
        return conv_norm(x, f"{p}.conv2", None, 1, False)

    # ------------------------------------------------------------------------------------------------ program
    def _compile(self, T: int, H: int, W: int) -> _Prog:
        if H % 8 or W % 8 or H < 128 or W < 128:
            raise _capi.VsrError("RAFT needs frames of at least 128 x 128 with sides divisible by 8 (the reference returns NaN flows below)")
        rt, w = self._rt, self.w
        prog = _Prog()
        prog.eager, prog.iter, prog.tail = [], [], []
        N = T - 1                     # pairs per direction
        h, wd = H // 8, W // 8
        hw = h * wd

        def new(c, hh, ww, n=T):
            cp = _r(c, 64)
            return _Tensor(rt.alloc(n * hh * ww * cp * 2), c, hh, ww, cp, n=n)

        prog.x = _Tensor(rt.alloc(T * H * W * 8 * 2), 3, H, W, 8, n=T)
        F = self._encoder(prog, new, prog.x, "fnet", "instance")
        Cx = self._encoder(prog, new, prog.x, "cnet", "batch")
        img = lambda t, k: t.ptr + k * hw * t.cp * 2         # noqa: E731  (image k of a [T,h,w,cp] tensor)

        # per-direction state (pairs batched on the image axis)
        dims = [(h, wd)]
        for _ in range(3):
            dims.append((dims[-1][0] // 2, dims[-1][1] // 2))
        pitches = [_r(a * b, 8) for a, b in dims]
        prog.dirs = []
        for d in (0, 1):
            st = _Prog()
            st.levels = [rt.alloc(N * hw * pitches[l] * 2) for l in range(4)]
            st.hx, st.rhx = new(384, h, wd, N), new(384, h, wd, N)
            st.flow32 = rt.alloc(N * hw * 2 * 4)
            st.flow16 = _Tensor(rt.alloc(N * hw * 8 * 2), 2, h, wd, 8, n=N)
            st.out32 = rt.alloc(N * 2 * H * W * 4)
            prog.dirs.append(st)
            for k in range(N):
                a, b = (k, k + 1) if d == 0 else (k + 1, k)
                prog.eager.append(lambda st=st, a=a, b=b, k=k: rt.corr_volume(img(F, a), img(F, b), h, wd, 256, st.levels[0] + k * hw * pitches[0] * 2, pitches[0]))
            for l in range(3):
                prog.eager.append(lambda st=st, l=l: rt.corr_pool(st.levels[l], N * hw, dims[l][0], dims[l][1], pitches[l], st.levels[l + 1], pitches[l + 1]))
            first = 0 if d == 0 else 1
            ctx = _Tensor(img(Cx, first), 256, h, wd, Cx.cp, n=N)     # images first .. first+N-1 are contiguous
            prog.eager.append(lambda st=st, ctx=ctx: rt.context_split(ctx, _view(st.hx, 0, 128), _view(st.hx, 128, 128)))
            prog.eager.append(lambda st=st: rt.copy_channels(_view(st.hx, 128, 128), st.rhx, 128, 128))

        # update block (shared weights), tensors shared by both directions (they run one after the other)
        P = "update_block"
        corr = new(324, h, wd, N)                 # pitch 384
        cor1, cf, fl1 = new(256, h, wd, N), new(256, h, wd, N), new(128, h, wd, N)
        z, r, q = new(128, h, wd, N), new(128, h, wd, N), new(128, h, wd, N)
        fh1, delta = new(256, h, wd, N), new(8, h, wd, N)
        m1, mask = new(256, h, wd, N), new(576, h, wd, N)
        L = lambda name, cin_pitch, pad=(0, 0), scale=1.0: self._conv((P, name), w[f"{P}.{name}.weight"] * np.float32(scale),   # noqa: E731
                                                                    w[f"{P}.{name}.bias"] * np.float32(scale), cin_pitch, 1, pad)
        lc1, lc2 = L("encoder.convc1", corr.cp), L("encoder.convc2", cor1.cp, (1, 1))
        lf1, lf2 = L("encoder.convf1", 8, (3, 3)), L("encoder.convf2", fl1.cp, (1, 1))
        lcv = L("encoder.conv", cf.cp, (1, 1))
        gz = [L(f"gru.convz{s}", 384, pad) for s, pad in (("1", (0, 2)), ("2", (2, 0)))]
        gr = [L(f"gru.convr{s}", 384, pad) for s, pad in (("1", (0, 2)), ("2", (2, 0)))]
        gq = [L(f"gru.convq{s}", 384, pad) for s, pad in (("1", (0, 2)), ("2", (2, 0)))]
        lh1, lh2 = L("flow_head.conv1", 384, (1, 1)), L("flow_head.conv2", fh1.cp, (1, 1))
        lm1, lm2 = L("mask.0", 384, (1, 1)), L("mask.2", m1.cp, (0, 0), 0.25)        # 0.25 * mask (update.py:137)

        def iteration(st):
            steps = []
            run = steps.append
            levels = [(st.levels[l], dims[l][0], dims[l][1], pitches[l]) for l in range(4)]
            hview = _view(st.hx, 0, 128)
            run(lambda: rt.corr_lookup(levels, st.flow32, h, wd, N * hw, corr))
            run(lambda: rt.conv_ex(lc1, corr, cor1, 1))
            run(lambda: rt.conv_ex(lc2, cor1, cf, 1, 0))                          # cor -> cf[:, 0:192]
            run(lambda: rt.conv_ex(lf1, st.flow16, fl1, 1))
            run(lambda: rt.conv_ex(lf2, fl1, cf, 1, 192))                         # flo -> cf[:, 192:256]
            run(lambda: rt.conv_ex(lcv, cf, st.hx, 1, 256))                       # out (126 + 2 zero) -> hx[:, 256:384]
            run(lambda: rt.copy_channels(_view(st.hx, 256, 128), st.rhx, 256, 128))
            run(lambda: rt.flow_update(st.flow32, None, st.flow16, st.hx, st.rhx, 382, False))   # cat([out, flow])
            for s in (0, 1):                                                       # SepConvGRU: 1x5 then 5x1
                run(lambda s=s: rt.conv_ex(gz[s], st.hx, z, 0))
                run(lambda s=s: rt.conv_ex(gr[s], st.hx, r, 0))
                run(lambda: rt.gru_rh(r, hview, _view(st.rhx, 0, 128)))
                run(lambda s=s: rt.conv_ex(gq[s], st.rhx, q, 0))
                run(lambda: rt.gru_update(z, q, hview))
            run(lambda: rt.conv_ex(lh1, hview, fh1, 1))
            run(lambda: rt.conv_ex(lh2, fh1, delta, 0))
            run(lambda: rt.flow_update(st.flow32, delta, st.flow16, None, None, 0, True))        # coords1 += delta
            return steps

        for st in prog.dirs:
            st.steps = iteration(st)
            hview = _view(st.hx, 0, 128)
            st.tail = [lambda hview=hview: rt.conv_ex(lm1, hview, m1, 1), lambda: rt.conv_ex(lm2, m1, mask, 0),
                       lambda st=st: rt.convex_upsample(st.flow32, mask, N, h, wd, st.out32)]
            st.graph = None
        prog.N, prog.h, prog.w, prog.hw = N, h, wd, hw
        return prog

    def __call__(self, frames_bgr: Sequence[np.ndarray], iters: int = ITERS, dst=None):
        """-> (forward flows t -> t+1, backward flows t+1 -> t), each float32 [T-1, 2, H, W] (x, y components).
        With `dst = (device pointer forward, device pointer backward)` the flows stay on the device: they are copied there (fp32
        [T-1,2,H,W] each) on the runtime's stream and nothing is downloaded (the caller hands them to the flow-completion network)."""
        frames = [np.ascontiguousarray(f, np.uint8) for f in frames_bgr]
        T = len(frames)
        if T < 2:
            raise ValueError("need at least two frames")
        H, W = frames[0].shape[:2]
        if any(f.shape != (H, W, 3) for f in frames):
            raise ValueError("frames must share one [H,W,3] shape")
        prog = self._progs.get((T, H, W))
        if prog is None:
            prog = self._progs[(T, H, W)] = self._compile(T, H, W)
        rt = self._rt
        rt.frames(frames, prog.x)
        for st in prog.eager:
            st()
        out = []
        for st in prog.dirs:
            rt.zero(st.flow32, prog.N * prog.hw * 8)
            rt.zero(st.flow16.ptr, prog.N * prog.hw * 16)
            for s in st.steps:                    # iteration 1 eagerly (every buffer exists afterwards), the rest as a graph
                s()
            if iters > 1:
                if st.graph is None:
                    rt.capture_begin()
                    try:
                        for s in st.steps:
                            s()
                    finally:
                        st.graph = rt.capture_end()
                for _ in range(iters - 1):
                    rt.graph_launch(st.graph)
            for s in st.tail:
                s()
            if dst is None:
                out.append(rt.download_f32(st.out32, (prog.N, 2, H, W)))
            else:
                rt.copy_bytes(st.out32, dst[len(out)], prog.N * 2 * H * W * 4)
                out.append(None)
        if rt.overflow():
            raise _capi.VsrError("RAFT activations left the fp16 range")
        return out[0], out[1]


__all__ = ["RaftFlow", "load_raft_weights"]
