"""ProPainter's InpaintGenerator on the device runtime.

STATUS: equal to the oracle (oracle/propainter_gen_oracle.py, pinned to the reference's frames) on the CPU stand-in of the runtime
(tests/test_generator_cpu.py: stage taps and final u8 predictions); kernels compile for sm_100a; NOT yet run on a B200.

Mirrors video/model/propainter.py:196-235 (Encoder: grouped convs over [x0 | out] re-interleaved per group — laid out with channel
copies, each group a tensor-core conv over a channel-slice view), :321-359 (inputs, 1/4 flows and masks) and :75-193 with
learnable=True (flow-guided deformable alignment: one fused kernel builds the 261-channel condition tensor = current features,
bilinear-warped propagated features, flow, consistency bit, masks; `deform_cols` + a K = 1152 GEMM is the deformable conv).
Back half (sparse_transformer.py): soft split = tap-major unfold + a K = 6272 GEMM (the linear layers around unfold / fold have their
weights permuted on the host from torch's channel-major c*49+k order to k*C+c); per block: layer norm, zero padding to 5x9 windows
BEFORE the q/k/v projections (padded tokens carry the biases, like the reference), a depthwise 4x4 pooling + the same projections
for the pooled tokens, ONE attention kernel that resolves own / rolled / pooled keys by index arithmetic (masked windows: every 2nd
frame, all keys; unmasked windows: the 45 tokens of the same frame), projection, residual; the fusion feed-forward as fc1 -> fold
with overlap normalisation -> unfold + GELU -> fc2.  The first version of the attention runs on CUDA cores (online softmax).
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

    return {k: (v.float().numpy() if v.is_floating_point() else v.numpy()) for k, v in torch.load(compact_checkpoint(path_or_dict), map_location="cpu").items()}


def compact_checkpoint(path) -> str:
    """`ProPainter.pth`, or next to it `ProPainter.f16.pth` (tools/stage_weights.py: the same state dict with its matrices / conv kernels
    stored in fp16 — what the tensor cores multiply with anyway; vectors stay fp32) when only that one was shipped (the gpurun snapshot is
    capped at 512 MiB)."""
    import os

    path = str(path)
    alt = path[:-4] + ".f16.pth" if path.endswith(".pth") else path
    return path if os.path.exists(path) or not os.path.exists(alt) else alt


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

    def unfold7s3(self, x, out, gelu=False):
        _capi.check(self.L.vsr_rt_unfold7s3(self.h, x.ptr, x.n, x.h, x.w, x.cp, out.ptr, out.cp, 1 if gelu else 0))

    def fold7s3(self, tok, out, channels, normalise):
        _capi.check(self.L.vsr_rt_fold7s3(self.h, tok.ptr, out.n, out.h, out.w, channels, tok.cp, 1 if normalise else 0, out.ptr))

    def layernorm(self, x, gamma, beta, out):
        _capi.check(self.L.vsr_rt_layernorm(self.h, x.ptr, x.pixels, x.cp, gamma, beta, out.ptr))

    def pool4(self, x, w_dev, b_dev, out):
        _capi.check(self.L.vsr_rt_pool4(self.h, x.ptr, x.n, x.h, x.w, x.cp, w_dev, b_dev, out.ptr))

    def window_attention(self, q, k, v, kp, vp, valid_dev, n_valid, tind_dev, n_tind, masked_dev, out):
        _capi.check(self.L.vsr_rt_window_attention(self.h, q.ptr, k.ptr, v.ptr, kp.ptr, vp.ptr, q.n, q.h, q.w, q.cp, kp.h, kp.w, valid_dev, n_valid, tind_dev,
                                                   n_tind, masked_dev, out.ptr))

    def pred_to_rgb8(self, x) -> np.ndarray:
        import ctypes as C

        out = np.empty((x.n, x.h, x.w, 3), np.uint8)
        _capi.check(self.L.vsr_rt_pred_to_rgb8(self.h, x.ptr, x.cp, x.pixels, _capi.ptr(out, C.c_uint8)))
        return out


class Generator:
    def __init__(self, weights, device="cuda:0", runtime=None):
        self.w = load_generator_weights(weights)
        self._rt = runtime if runtime is not None else _GenRuntime(device)
        self._layers: Dict[tuple, int] = {}
        self._arena = _Arena(self._rt)
        self._consts: Dict[tuple, int] = {}

    def _const(self, key, arr) -> int:
        """device copy of a small constant array, uploaded once"""
        if key not in self._consts:
            arr = np.ascontiguousarray(arr)
            self._consts[key] = self._rt.upload_f32(arr) if arr.dtype == np.float32 else self._rt.upload_bytes(arr)
        return self._consts[key]

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

        ids_dev = self._const(("ids", tuple(ids)), np.asarray(list(ids), np.int32))
        flow_ids = self._const(("ids", tuple(ids[:l_t - 1])), np.asarray(list(ids[:l_t - 1]), np.int32))   # local flows: neighbour ids[:-1]
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


    # ------------------------------------------------------------------------------------------------ back half
    @staticmethod
    def token_mask(mask: np.ndarray) -> np.ndarray:
        """mask_pool_l of one local frame (:350-358): nearest 1/4 of the input mask, MaxPool2d(7, 3, 3) -> [fh, fw] in {0, 1}."""
        ds = (np.asarray(mask)[::4, ::4] > 0).astype(np.uint8)
        h, w = ds.shape
        fh, fw = (h + 6 - 7) // 3 + 1, (w + 6 - 7) // 3 + 1
        p = np.pad(ds, 3)
        out = np.zeros((fh, fw), np.uint8)
        for ty in range(fh):
            for tx in range(fw):
                out[ty, tx] = p[3 * ty:3 * ty + 7, 3 * tx:3 * tx + 7].max()
        return out

    def _lin(self, key, weight, bias, x: _Tensor, y: _Tensor):
        """nn.Linear over the channel axis of a token map = 1x1 conv"""
        self._rt.conv_ex(self._conv(key, np.ascontiguousarray(weight[:, :, None, None], np.float32), bias, x.cp, 1, 0), x, y, 0)

    def transform_and_decode(self, enc: _Tensor, l_t: int, mask: np.ndarray, H: int, W: int) -> np.ndarray:
        """enc: [n,h,w,128] after feature propagation; mask: the host u8 mask of the strip -> u8 predictions [l_t,H,W,3] (RGB, before the
        mask composite of P7)."""
        rt, w = self._rt, self.w
        n, h, wd = enc.n, enc.h, enc.w
        self._arena.begin(("back", n, l_t, H, W))
        alloc = self._arena.alloc

        def new(c, hh, ww, k=n, cp=None):
            cp = cp or _r(c, 64)
            return _Tensor(alloc(k * hh * ww * cp * 2), c, hh, ww, cp, n=k)

        fh, fw = (h + 6 - 7) // 3 + 1, (wd + 6 - 7) // 3 + 1
        Hn, Wn = -(-fh // 5) * 5, -(-fw // 9) * 9
        tm = np.zeros((Hn, Wn), np.uint8)
        tm[:fh, :fw] = self.token_mask(mask)
        win_masked = tm.reshape(Hn // 5, 5, Wn // 9, 9).max((1, 3)).reshape(-1).astype(np.int32)     # the same for every local frame
        masked_dev = self._const(("masked", win_masked.tobytes()), win_masked)
        t_inds = [np.arange(i, n, 2, dtype=np.int32) for i in range(2)]
        tind_dev = [self._const(("tind", n, i), t) for i, t in enumerate(t_inds)]

        def tapmajor_cols(wt, C):      # [out, C*49] (c*49 + k) -> [out, 49*C] (k*C + c)
            return np.ascontiguousarray(wt.reshape(wt.shape[0], C, 49).transpose(0, 2, 1).reshape(wt.shape[0], 49 * C))

        def tapmajor_rows(wt, b, C):   # rows c*49 + k -> k*C + c
            return (np.ascontiguousarray(wt.reshape(C, 49, wt.shape[1]).transpose(1, 0, 2).reshape(49 * C, wt.shape[1])),
                    np.ascontiguousarray(b.reshape(C, 49).T.reshape(-1)))

        # SoftSplit (:7-32)
        u = new(6272, fh, fw)
        rt.unfold7s3(enc, u)
        tok = new(512, fh, fw)
        self._lin("ss", tapmajor_cols(w["ss.embedding.weight"], 128), w["ss.embedding.bias"], u, tok)
        for i in range(8):
            p = f"transformers.transformer.{i}"
            g1, b1 = self._const((p, "g1"), w[f"{p}.norm1.weight"]), self._const((p, "b1"), w[f"{p}.norm1.bias"])
            y = new(512, fh, fw)
            rt.layernorm(tok, g1, b1, y)
            yp = y
            if (Hn, Wn) != (fh, fw):
                yp = new(512, Hn, Wn)
                rt.pad(y, yp, 0, 0, 0)                                             # zeros at the bottom / right BEFORE the projections
            q, k, v = new(512, Hn, Wn), new(512, Hn, Wn), new(512, Hn, Wn)
            a = f"{p}.attention"
            for name, dst in (("query", q), ("key", k), ("value", v)):
                self._lin((a, name), w[f"{a}.{name}.weight"], w[f"{a}.{name}.bias"], yp, dst)
            ph, pw = Hn // 4, Wn // 4
            pooled = new(512, ph, pw)
            rt.pool4(yp, self._const((a, "pw"), w[f"{a}.pool_layer.weight"].reshape(512, 16)), self._const((a, "pb"), w[f"{a}.pool_layer.bias"]), pooled)
            kp, vp = new(512, ph, pw), new(512, ph, pw)
            self._lin((a, "key"), w[f"{a}.key.weight"], w[f"{a}.key.bias"], pooled, kp)
            self._lin((a, "value"), w[f"{a}.value.weight"], w[f"{a}.value.bias"], pooled, vp)
            att = new(512, Hn, Wn)
            valid = np.asarray(w[f"{a}.valid_ind_rolled"], np.int32)
            rt.window_attention(q, k, v, kp, vp, self._const((a, "valid"), valid), int(valid.size), tind_dev[i % 2], int(t_inds[i % 2].size), masked_dev, att)
            ac = att
            if (Hn, Wn) != (fh, fw):
                ac = new(512, fh, fw)
                rt.pad(att, ac, 0, 0, 0)                                           # crop back
            pr = new(512, fh, fw)
            self._lin((a, "proj"), w[f"{a}.proj.weight"], w[f"{a}.proj.bias"], ac, pr)
            rt.elementwise(0, tok, pr, tok)
            g2, b2 = self._const((p, "g2"), w[f"{p}.norm2.weight"]), self._const((p, "b2"), w[f"{p}.norm2.bias"])
            y2 = new(512, fh, fw)
            rt.layernorm(tok, g2, b2, y2)
            w1, bb1 = tapmajor_rows(w[f"{p}.mlp.fc1.0.weight"], w[f"{p}.mlp.fc1.0.bias"], 40)
            f1 = new(1960, fh, fw)
            self._lin((p, "fc1"), w1, bb1, y2, f1)
            fm = new(40, h, wd, cp=40)
            rt.fold7s3(f1, fm, 40, True)
            f2 = new(1960, fh, fw)
            rt.unfold7s3(fm, f2, gelu=True)
            o = new(512, fh, fw)
            self._lin((p, "fc2"), tapmajor_cols(w[f"{p}.mlp.fc2.1.weight"], 40), w[f"{p}.mlp.fc2.1.bias"], f2, o)
            rt.elementwise(0, tok, o, tok)
        # SoftComp (:35-66) + residual
        ws, bs = tapmajor_rows(w["sc.embedding.weight"], w["sc.embedding.bias"], 128)
        z = new(6272, fh, fw)
        self._lin("sc", ws, bs, tok, z)
        zf = new(128, h, wd)
        rt.fold7s3(z, zf, 128, False)
        zc = new(128, h, wd)
        rt.conv_ex(self._conv("sc.bias_conv", w["sc.bias_conv.weight"], w["sc.bias_conv.bias"], zf.cp, 1, 1), zf, zc, 0)
        rt.elementwise(0, enc, zc, zc)
        # decoder on the local frames (:370-376)
        y = _Tensor(zc.ptr, 128, h, wd, zc.cp, n=l_t)
        for name, up in (("decoder.0.conv", True), ("decoder.2", False), ("decoder.4.conv", True)):
            if up:
                u2 = new(y.c, 2 * y.h, 2 * y.w, l_t)
                rt.upsample2x(y, u2)
                y = u2
            wt = w[f"{name}.weight"]
            o = new(wt.shape[0], y.h, y.w, l_t)
            rt.conv_ex(self._conv(name, wt, w[f"{name}.bias"], y.cp, 1, 1), y, o, 0)
            rt.leaky(o, 0.2)
            y = o
        o = new(8, y.h, y.w, l_t)
        rt.conv_ex(self._conv("decoder.6", w["decoder.6.weight"], w["decoder.6.bias"], y.cp, 1, 1), y, o, 0)
        if rt.overflow():
            raise _capi.VsrError("generator activations left the fp16 range")
        return rt.pred_to_rgb8(o)


__all__ = ["Generator", "load_generator_weights"]
