"""ProPainter's recurrent flow completion on the device runtime.

STATUS: like raft_flow.py — equal to the oracle (oracle/rfc_oracle.py, pinned to the reference's completed flows) on the CPU
stand-in of the runtime (tests/test_flow_completion_cpu.py); kernels compile for sm_100a; NOT yet run on a B200.

Mirrors `RecurrentFlowCompleteNet.forward_bidirect_flow` + `combine_flow` (video/model/recurrent_flow_completion.py:275-348) for one
mask shared by all frames (what `read_mask` produces for the pipeline).  Mapping:
* Conv3d (1,k,k) layers are 2-D convs over the frame batch; the first one's replicate padding is a pad launch + a direct conv;
* P3DBlock's temporal conv (3,1,1), dilation 2: `temporal_taps` lays frames t-2, t, t+2 side by side, a 1x1 conv finishes it;
* second-order deformable alignment: the offset / mask convs run on the conv kernel, `deform_cols` gathers the modulated
  bilinear samples (16 groups x 9 taps) and a 1x1 conv with K = 9*256 is the deformable conv itself;
* the propagation is sequential over frames, as in the reference; LeakyReLU is a separate in-place launch (the validated conv
  epilogues only know ReLU; folding it is a later optimisation).
"""
from typing import Dict

import numpy as np

from . import _capi
from .dbnet import _Tensor, _r
from .flow_propagation import _Arena, _PropRuntime, _image
from .lama_inpaint import _view


def load_rfc_weights(path_or_dict) -> Dict[str, np.ndarray]:
    if isinstance(path_or_dict, dict):
        return {k: np.asarray(v, np.float32) for k, v in path_or_dict.items()}
    import torch

    return {k: v.float().numpy() for k, v in torch.load(str(path_or_dict), map_location="cpu").items()}


class _RfcRuntime(_PropRuntime):
    def rfc_input(self, flow32, mask_u8, n, hh, ww, reverse, out):
        _capi.check(self.L.vsr_rt_rfc_input(self.h, flow32, mask_u8, n, hh, ww, 1 if reverse else 0, out.ptr))

    def pad_replicate(self, x, y, top, left):
        _capi.check(self.L.vsr_rt_pad_replicate(self.h, x.ptr, x.n, x.h, x.w, x.cp, y.ptr, y.h, y.w, top, left))

    def leaky(self, x, slope):
        _capi.check(self.L.vsr_rt_leaky_relu(self.h, x.ptr, x.pixels * x.cp, slope))

    def temporal_taps(self, x, y):
        _capi.check(self.L.vsr_rt_temporal_taps(self.h, x.ptr, x.n, x.h * x.w, x.cp, y.ptr, y.cp))

    def deform_cols(self, xa, ca, xb, c, groups, om, max_residue, flow32, cols):
        _capi.check(self.L.vsr_rt_deform_cols(self.h, xa.ptr, xa.cp, ca, xb.ptr if xb is not None else 0, xb.cp if xb is not None else 0, c, groups, om.ptr,
                                              om.cp, max_residue, flow32, xa.h, xa.w, xa.pixels, cols.ptr, cols.cp))

    def rfc_combine(self, pred, flow32, mask_u8, n, hh, ww, reverse, out32):
        _capi.check(self.L.vsr_rt_rfc_combine(self.h, pred.ptr, pred.cp, flow32, mask_u8, n, hh, ww, 1 if reverse else 0, out32))

    def upsample2x(self, x, y):
        _capi.check(self.L.vsr_rt_upsample2x_bilinear(self.h, x.ptr, x.n, x.h, x.w, x.cp, y.ptr))


class FlowCompletion:
    """complete(flows_f, flows_b, mask) on device buffers; `complete_host` for numpy in / out."""

    def __init__(self, weights, device="cuda:0", runtime=None):
        self.w = load_rfc_weights(weights)
        self._rt = runtime if runtime is not None else _RfcRuntime(device)
        self._layers: Dict[tuple, int] = {}
        self._arena = _Arena(self._rt)

    def __del__(self):
        rt = getattr(self, "_rt", None)
        if rt is not None:
            try:
                rt.close()
            except Exception:
                pass
            self._rt = None

    def _conv(self, key, weight, bias, cin_pitch, stride=1, pad=0, dil=1):
        k = (key, cin_pitch)
        if k not in self._layers:
            cout, cin, kh, kw = weight.shape
            if cout >= 8 and cout % 8:
                padn = _r(cout, 8) - cout
                weight = np.concatenate([weight, np.zeros((padn,) + weight.shape[1:], np.float32)])
                bias = np.concatenate([bias, np.zeros(padn, np.float32)])
            self._layers[k] = self._rt.conv_create(weight, bias, weight.shape[0], cin, cin_pitch, kh, kw, stride, pad, pad, dil, 1, False)
        return self._layers[k]

    def _network(self, flow32: int, mask_u8: int, N: int, H: int, W: int, reverse: bool, out32: int):
        """RecurrentFlowCompleteNet.forward on the (optionally time-reversed) masked flows, combined with the input flows."""
        if H % 8 or W % 8:
            raise _capi.VsrError("flow completion needs sides divisible by 8")
        rt, w = self._rt, self.w
        self._arena.begin((N, H, W, bool(reverse)))
        alloc = self._arena.alloc

        def new(c, hh, ww, n=N):
            cp = _r(c, 64)
            return _Tensor(alloc(n * hh * ww * cp * 2), c, hh, ww, cp, n=n)

        def conv(x, name, cout_hw, stride=1, pad=1, dil=1, slope=None, wkey=None, squeeze=True):
            wt = w[f"{name}.weight"]
            wt = wt[:, :, 0] if wt.ndim == 5 and squeeze else wt            # Conv3d (1,k,k) -> Conv2d
            y = new(wt.shape[0], *cout_hw)
            rt.conv_ex(self._conv(wkey or name, wt, w[f"{name}.bias"], x.cp, stride, pad, dil), x, y, 0)
            if slope is not None:
                rt.leaky(y, slope)
            return y

        def p3d(x, name, stride):
            """P3DBlock: spatial conv + LeakyReLU(0.2), temporal (3,1,1) dilation-2 conv; the encoder's LeakyReLU(0.2) follows."""
            y = conv(x, f"{name}.conv1.0", (x.h // stride, x.w // stride), stride, 1, 1, 0.2)
            taps = _Tensor(alloc(N * y.h * y.w * 3 * y.cp * 2), 3 * y.cp, y.h, y.w, 3 * y.cp, n=N)
            rt.temporal_taps(y, taps)
            w3 = w[f"{name}.conv2.0.weight"][:, :, :, 0, 0]                  # [Cout, C, 3]
            wt = np.zeros((w3.shape[0], 3 * y.cp, 1, 1), np.float32)
            for k in range(3):
                wt[:, k * y.cp: k * y.cp + w3.shape[1], 0, 0] = w3[:, :, k]
            z = new(w3.shape[0], y.h, y.w)
            rt.conv_ex(self._conv((name, "t"), wt, w[f"{name}.conv2.0.bias"], taps.cp, 1, 0), taps, z, 0)
            rt.leaky(z, 0.2)
            return z

        x0 = _Tensor(alloc(N * H * W * 8 * 2), 3, H, W, 8, n=N)
        rt.rfc_input(flow32, mask_u8, N, H, W, reverse, x0)
        xp = _Tensor(alloc(N * (H + 4) * (W + 4) * 8 * 2), 3, H + 4, W + 4, 8, n=N)
        rt.pad_replicate(x0, xp, 2, 2)
        x = conv(xp, "downsample.0", (H // 2, W // 2), 2, 0, 1, 0.2)                              # 5x5 stride 2 on the replicate-padded input
        e1 = p3d(p3d(x, "encoder1.0", 1), "encoder1.2", 2)                                         # /4, 64 ch
        e2 = p3d(p3d(e1, "encoder2.0", 1), "encoder2.2", 2)                                        # /8, 128 ch
        m = e2
        for i, d in ((0, 3), (2, 2), (4, 1)):
            m = conv(m, f"mid_dilation.{i}", (m.h, m.w), 1, d, d, 0.2)
        f = self._propagate(m, new)
        d2 = conv(f, "decoder2.0", (f.h, f.w), 1, 1, 1, 0.2)
        u = new(d2.c, 2 * d2.h, 2 * d2.w)
        rt.upsample2x(d2, u)
        d2 = conv(u, "decoder2.2.conv", (u.h, u.w), 1, 1, 1, 0.2)
        rt.elementwise(0, d2, e1, d2)                                                               # + feat_e1
        d1 = conv(d2, "decoder1.0", (d2.h, d2.w), 1, 1, 1, 0.2)
        u = new(d1.c, 2 * d1.h, 2 * d1.w)
        rt.upsample2x(d1, u)
        d1 = conv(u, "decoder1.2.conv", (u.h, u.w), 1, 1, 1, 0.2)
        up = conv(d1, "upsample.0", (d1.h, d1.w), 1, 1, 1, 0.2)
        u = new(up.c, 2 * up.h, 2 * up.w)
        rt.upsample2x(up, u)
        flow = conv(u, "upsample.2.conv", (u.h, u.w), 1, 1, 1)                                       # 2 channels (direct kernel)
        rt.rfc_combine(flow, flow32, mask_u8, N, H, W, reverse, out32)

    def _propagate(self, x: _Tensor, new) -> _Tensor:
        """BidirectionalPropagation.forward (:70-128) on [N,h,w,128]."""
        rt, w, N = self._rt, self.w, x.n
        p = "feat_prop_module"
        h, wd, c = x.h, x.w, 128
        one = lambda ch: new(ch, h, wd, 1)                                    # noqa: E731
        zeros = one(c)
        feats = {"backward_": [None] * N, "forward_": [None] * N}
        for name in ("backward_", "forward_"):
            order = list(range(N - 1, -1, -1)) if name == "backward_" else list(range(N))
            q = f"{p}.deform_align.{name}"
            wdc = w[f"{q}.weight"]                                                # [128, 256, 3, 3] -> 1x1 over cols [k*256 + c]
            wcols = np.ascontiguousarray(wdc.transpose(0, 2, 3, 1).reshape(128, 9 * 256, 1, 1))
            prop, hist = zeros, []
            for i, idx in enumerate(order):
                cur = _image(x, idx)
                if i > 0:
                    n2 = hist[-2] if i > 1 else zeros
                    cond = one(384)
                    for off, t in ((0, prop), (128, cur), (256, n2)):
                        rt.copy_channels(t, cond, off, 128)
                    o = cond
                    for j in (0, 2, 4):
                        y = one(128)
                        rt.conv_ex(self._conv((q, j), w[f"{q}.conv_offset.{j}.weight"], w[f"{q}.conv_offset.{j}.bias"], o.cp, 1, 1), o, y, 0)
                        rt.leaky(y, 0.1)
                        o = y
                    om = one(432)                                                # pitch 448
                    rt.conv_ex(self._conv((q, 6), w[f"{q}.conv_offset.6.weight"], w[f"{q}.conv_offset.6.bias"], o.cp, 1, 1), o, om, 0)
                    cols = one(9 * 256)
                    rt.deform_cols(prop, 128, n2, 256, 16, om, 5.0, 0, cols)
                    aligned = one(128)
                    rt.conv_ex(self._conv((q, "dc"), wcols, w[f"{q}.bias"], cols.cp, 1, 0), cols, aligned, 0)
                    prop = aligned
                others = [feats[k][idx] for k in ("backward_", "forward_") if k != name and feats[k][idx] is not None]
                parts = [cur] + others + [prop]
                cat = one(128 * len(parts))
                for j, t in enumerate(parts):
                    rt.copy_channels(t, cat, 128 * j, 128)
                b = f"{p}.backbone.{name}"
                y = one(128)
                rt.conv_ex(self._conv((b, 0), w[f"{b}.0.weight"], w[f"{b}.0.bias"], cat.cp, 1, 1), cat, y, 0)
                rt.leaky(y, 0.1)
                z = one(128)
                rt.conv_ex(self._conv((b, 2), w[f"{b}.2.weight"], w[f"{b}.2.bias"], y.cp, 1, 1), y, z, 0)
                nxt = one(128)
                rt.elementwise(0, prop, z, nxt)
                prop = nxt
                hist.append(prop)
                feats[name][idx] = prop
        out = new(128, h, wd, N)
        lf = self._conv((p, "fusion"), w[f"{p}.fusion.weight"], w[f"{p}.fusion.bias"], 256, 1, 0)
        for i in range(N):
            cat = one(256)
            rt.copy_channels(feats["backward_"][i], cat, 0, 128)
            rt.copy_channels(feats["forward_"][i], cat, 128, 128)
            rt.conv_ex(lf, cat, _image(out, i), 0)
        rt.elementwise(0, out, x, out)
        return out

    def complete(self, flows_f: int, flows_b: int, mask_u8: int, N: int, H: int, W: int):
        """device fp32 [N,2,H,W] flows + device u8 [H,W] flow mask -> two new device fp32 [N,2,H,W] buffers (completed flows)."""
        rt = self._rt
        self._arena.begin(("out", N, H, W))
        out_f, out_b = self._arena.alloc(N * 2 * H * W * 4), self._arena.alloc(N * 2 * H * W * 4)
        self._network(flows_f, mask_u8, N, H, W, False, out_f)
        self._network(flows_b, mask_u8, N, H, W, True, out_b)
        if rt.overflow():
            raise _capi.VsrError("flow completion activations left the fp16 range")
        return out_f, out_b

    def complete_host(self, flows_f: np.ndarray, flows_b: np.ndarray, mask: np.ndarray):
        rt = self._rt
        N, _, H, W = flows_f.shape
        of, ob = self.complete(rt.upload_bytes(np.ascontiguousarray(flows_f, np.float32)), rt.upload_bytes(np.ascontiguousarray(flows_b, np.float32)),
                               rt.upload_bytes(np.ascontiguousarray(mask, np.uint8)), N, H, W)
        return rt.download_f32(of, (N, 2, H, W)), rt.download_f32(ob, (N, 2, H, W))


__all__ = ["FlowCompletion", "load_rfc_weights"]
