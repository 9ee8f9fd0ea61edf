"""ProPainter's InpaintGenerator on the device runtime — FRONT HALF: encoder and learnable feature propagation.

STATUS: equal to the oracle's stage taps (oracle/propainter_gen_oracle.py, pinned to the reference's frames) on the CPU stand-in of
the runtime (tests/test_generator_cpu.py); kernels compile for sm_100a; NOT yet run on a B200; the back half (soft split, the 8
sparse-window transformer blocks, soft composition, decoder) is not written yet.

Mirrors video/model/propainter.py:196-235 (Encoder: grouped convs over [x0 | out] re-interleaved per group — laid out with channel
copies, each group a tensor-core conv over a channel-slice view), :321-359 (inputs, 1/4 flows and masks) and :75-193 with
learnable=True (flow-guided deformable alignment: one fused kernel builds the 261-channel condition tensor = current features,
bilinear-warped propagated features, flow, consistency bit, masks; `deform_cols` + a K = 1152 GEMM is the deformable conv).
"""
from typing import Dict, List, Sequence

import numpy as np

from . import _capi
from .dbnet import _Tensor, _r
from .flow_completion import _RfcRuntime
from .flow_propagation import _Arena, _image
from .lama_inpaint import _view


def load_generator_weights(path_or_dict) -> Dict[str, np.ndarray]:
    if isinstance(path_or_dict, dict):
        return {k: np.asarray(v) for k, v in path_or_dict.items()}
    import torch

    return {k: (v.float().numpy() if v.is_floating_point() else v.numpy()) for k, v in torch.load(str(path_or_dict), map_location="cpu").items()}


class _GenRuntime(_RfcRuntime):
    def gen_input(self, state, mask_u8, ids_dev, n, out):
        _capi.check(self.L.vsr_rt_gen_input(self.h, state.ptr, mask_u8, ids_dev, n, state.h, state.w, out.ptr))

    def flow_down4(self, flow32, ids_dev, n, hh, ww, out32):
        _capi.check(self.L.vsr_rt_flow_down4(self.h, flow32, ids_dev, n, hh, ww, out32))

    def prop_masks(self, gen_in, out):
        _capi.check(self.L.vsr_rt_prop_masks(self.h, gen_in.ptr, gen_in.n, gen_in.h, gen_in.w, out.ptr))

    def featprop_cond(self, prop, cur, flow_prop, flow_check, masks, cond):
        _capi.check(self.L.vsr_rt_featprop_cond(self.h, prop.ptr, cur.ptr, cur.cp, flow_prop, flow_check, masks.ptr, cur.h, cur.w, cond.ptr, cond.cp))

    def write_extra(self, src, dst, coff, nch):
        _capi.check(self.L.vsr_rt_write_extra(self.h, src.ptr, dst.ptr, dst.cp, coff, nch, dst.pixels))


class Generator:
    def __init__(self, weights, device="cuda:0", runtime=None):
        self.w = load_generator_weights(weights)
        self._rt = runtime if runtime is not None else _GenRuntime(device)
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

    def _conv(self, key, weight, bias, cin_pitch, stride=1, pad=1):
        k = (key, cin_pitch)
        if k not in self._layers:
            cout, cin, kh, kw = weight.shape
            self._layers[k] = self._rt.conv_create(np.ascontiguousarray(weight, np.float32), np.ascontiguousarray(bias, np.float32), cout, cin, cin_pitch, kh, kw,
                                                   stride, pad, pad, 1, 1, False)
        return self._layers[k]

    # ------------------------------------------------------------------------------------------------ encoder
    def _encoder(self, x: _Tensor, new) -> _Tensor:
        """Encoder.forward (:222-235) on [n,H,W,8] (5 real channels) -> [n,H/4,W/4,128]."""
        rt, w = self._rt, self.w
        out, x0 = x, None
        plan = ((0, 2, 1), (2, 1, 1), (4, 2, 1), (6, 1, 1), (8, 1, 1), (10, 1, 2), (12, 1, 4), (14, 1, 8), (16, 1, 1))
        for i, stride, g in plan:
            wt, b = w[f"encoder.layers.{i}.weight"], w[f"encoder.layers.{i}.bias"]
            if i == 8:
                x0 = out
            if i > 8:                                   # [x0 | out] re-interleaved per group (:229-233), then a grouped conv
                a, c = x0.c // g, out.c // g
                cat = new(_r(g * (a + c) + 64, 64), out.h, out.w)      # one spare 64-channel box: the last group's K padding stays inside the pixel
                for j in range(g):
                    rt.copy_channels(_view(x0, j * a, a), cat, j * (a + c), a)
                    rt.copy_channels(_view(out, j * c, c), cat, j * (a + c) + a, c)
                y = new(wt.shape[0], out.h, out.w)
                cg = wt.shape[0] // g
                for j in range(g):
                    lid = self._conv((i, j), wt[j * cg:(j + 1) * cg], b[j * cg:(j + 1) * cg], cat.cp, 1, 1)
                    rt.conv_ex(lid, _view(cat, j * (a + c), a + c), y, 0, j * cg)
            else:
                y = new(wt.shape[0], out.h // stride, out.w // stride)
                rt.conv_ex(self._conv((i, 0), wt, b, out.cp, stride, 1), out, y, 0)
            rt.leaky(y, 0.2)
            out = y
        return out

    # ------------------------------------------------------------------------------------------------ feature propagation
    def _feature_propagation(self, x: _Tensor, flows_f: int, flows_b: int, masks: _Tensor, new) -> _Tensor:
        """BidirectionalPropagation(128, learnable=True).forward (:107-193) on the local frames [l,h,w,128]; flows fp32 [(l-1)*h*w][2]
        at feature resolution; masks fp16 [l*h*w][8] = (m_in, m_updated)."""
        rt, w, L, h, wd = self._rt, self.w, x.n, x.h, x.w
        p = "feat_prop_module"
        one = lambda c: new(c, h, wd, 1)                                   # noqa: E731
        fl = lambda base, i: base + i * h * wd * 2 * 4                     # noqa: E731
        feats: List[_Tensor] = [_image(x, i) for i in range(L)]
        outs = {}
        for name in ("backward_1", "forward_1"):
            order = list(range(L - 1, -1, -1)) if name == "backward_1" else list(range(L))
            q = f"{p}.deform_align.{name}"
            wcols = np.ascontiguousarray(w[f"{q}.weight"].transpose(0, 2, 3, 1).reshape(128, 9 * 128, 1, 1))
            res: List[_Tensor] = [None] * L
            prop = None
            for n, idx in enumerate(order):
                cur, mcur = feats[idx], _image(masks, idx)
                if n == 0:
                    prop = cur
                else:
                    fi = idx if name == "backward_1" else idx - 1
                    fprop, fcheck = (fl(flows_f, fi), fl(flows_b, fi)) if name == "backward_1" else (fl(flows_b, fi), fl(flows_f, fi))
                    cond = one(320)                                              # 261 used
                    rt.featprop_cond(prop, cur, fprop, fcheck, mcur, cond)
                    o = cond
                    for j in (0, 2, 4):
                        y = one(128)
                        wt = w[f"{q}.conv_offset.{j}.weight"]
                        rt.conv_ex(self._conv((q, j), wt, w[f"{q}.conv_offset.{j}.bias"], o.cp, 1, 1), o, y, 0)
                        rt.leaky(y, 0.1)
                        o = y
                    om = one(432)
                    rt.conv_ex(self._conv((q, 6), w[f"{q}.conv_offset.6.weight"], w[f"{q}.conv_offset.6.bias"], o.cp, 1, 1), o, om, 0)
                    cols = one(9 * 128)
                    rt.deform_cols(prop, 128, None, 128, 16, om, 3.0, fprop, cols)
                    aligned = one(128)
                    rt.conv_ex(self._conv((q, "dc"), wcols, w[f"{q}.bias"], cols.cp, 1, 0), cols, aligned, 0)
                    prop = aligned
                cat = one(320)                                                   # [cur | prop | masks(2)] = 258
                rt.copy_channels(cur, cat, 0, 128)
                rt.copy_channels(prop, cat, 128, 128)
                rt.write_extra(mcur, cat, 256, 2)
                b = f"{p}.backbone.{name}"
                y = one(128)
                rt.conv_ex(self._conv((b, 0), w[f"{b}.0.weight"], w[f"{b}.0.bias"], cat.cp, 1, 1), cat, y, 0)
                rt.leaky(y, 0.2)
                z = one(128)
                rt.conv_ex(self._conv((b, 2), w[f"{b}.2.weight"], w[f"{b}.2.bias"], y.cp, 1, 1), y, z, 0)
                nxt = one(128)
                rt.elementwise(0, prop, z, nxt)
                prop = nxt
                res[idx] = prop
            outs[name] = res
            feats = res                                                          # the forward pass runs over the backward pass's output
        out = new(128, h, wd, L)
        for i in range(L):
            cat = one(320)
            rt.copy_channels(outs["backward_1"][i], cat, 0, 128)
            rt.copy_channels(outs["forward_1"][i], cat, 128, 128)
            rt.write_extra(_image(masks, i), cat, 256, 2)
            y = one(128)
            rt.conv_ex(self._conv((p, "fuse0"), w[f"{p}.fuse.0.weight"], w[f"{p}.fuse.0.bias"], cat.cp, 1, 1), cat, y, 0)
            rt.leaky(y, 0.2)
            rt.conv_ex(self._conv((p, "fuse2"), w[f"{p}.fuse.2.weight"], w[f"{p}.fuse.2.bias"], y.cp, 1, 1), y, _image(out, i), 0)
        rt.elementwise(0, out, x, out)
        return out

    # ------------------------------------------------------------------------------------------------ front half
    def encode_and_propagate(self, state: _Tensor, mask_u8: int, ids: Sequence[int], flows_f: int, flows_b: int, l_t: int):
        """state: P5's output [T,H,W,8]; ids = neighbour ids + reference ids; flows: device fp32 [T-1,2,H,W] (completed flows).
        -> (features [len(ids), H/4, W/4, 128] with the first l_t frames propagated, masks tensor (m_in, m_updated) of the local frames)."""
        rt = self._rt
        T, H, W = state.n, state.h, state.w
        n = len(ids)
        if H % 8 or W % 8:
            raise _capi.VsrError("the generator needs sides divisible by 8")
        self._arena.begin(("front", n, l_t, H, W))
        alloc = self._arena.alloc

        def new(c, hh, ww, k=n):
            cp = _r(c, 64)
            return _Tensor(alloc(k * hh * ww * cp * 2), c, hh, ww, cp, n=k)

        ids_dev = rt.upload_bytes(np.asarray(list(ids), np.int32))
        flow_ids = rt.upload_bytes(np.asarray(list(ids[:l_t - 1]), np.int32))     # the local flows are those of neighbour ids[:-1]
        gin = _Tensor(alloc(n * H * W * 8 * 2), 5, H, W, 8, n=n)
        rt.gen_input(state, mask_u8, ids_dev, n, gin)
        enc = self._encoder(gin, new)
        h, wd = enc.h, enc.w
        df, db = alloc((l_t - 1) * h * wd * 2 * 4), alloc((l_t - 1) * h * wd * 2 * 4)
        rt.flow_down4(flows_f, flow_ids, l_t - 1, H, W, df)
        rt.flow_down4(flows_b, flow_ids, l_t - 1, H, W, db)
        masks = _Tensor(alloc(l_t * h * wd * 8 * 2), 2, h, wd, 8, n=l_t)
        local_in = _Tensor(gin.ptr, 5, H, W, 8, n=l_t)
        rt.prop_masks(local_in, masks)
        local = _Tensor(enc.ptr, 128, h, wd, enc.cp, n=l_t)
        prop = self._feature_propagation(local, df, db, masks, new)
        rt.copy_channels(prop, local, 0, 128)                                     # enc_feat = cat(local_feat, ref_feat)
        return enc, masks


__all__ = ["Generator", "load_generator_weights"]
