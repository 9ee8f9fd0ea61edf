"""CPU oracle for ProPainter's generator and the whole `PropainterInpaint.inpaint` chain — TEST
INFRASTRUCTURE ONLY.

Restates, functionally over `ProPainter.pth`:
  InpaintGenerator.forward                      video/model/propainter.py:321-378
  Encoder (grouped convs over [x0 | out])       propainter.py:196-235
  BidirectionalPropagation(learnable=True) with flow-guided DeformableAlignment   propainter.py:36-72,75-193
  SoftSplit / SoftComp / FusionFeedForward      video/model/modules/sparse_transformer.py:7-112
  SparseWindowAttention (5x9 windows, 4 heads, rolled + 4x4-pooled keys, masked windows attend every 2nd frame)   :127-283
  TemporalSparseTransformer(Block)              :286-344
and `inpaint` (propainter_inpaint.py:192-361) for clips up to sub_video_length frames by chaining the other oracles
(P2 read_mask -> P3 RAFT -> P4 flow completion -> P5 image propagation -> P6 -> P7 composite).  The long-video branches
(:219-236 RAFT clips aside, :244-268, :283-306 sub-video chunking) are not restated.
Parity: PINNED against tests/golden/propainter_real.npz (`comp`, `call`: frames of the unmodified reference).
"""
import math
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from oracle import propainter_oracle as P
from oracle.deform_conv import deform_conv2d

WINDOW, HEADS, POOL, DEPTH, T2T = (5, 9), 4, (4, 4), 8, dict(kernel_size=(7, 7), stride=(3, 3), padding=(3, 3))


def load_weights(path: str) -> Dict[str, torch.Tensor]:
    import os

    alt = path[:-4] + ".f16.pth"     # tools/stage_weights.py: the compact copy that travels to the GPU box (matrices stored fp16)
    if not os.path.exists(path) and os.path.exists(alt):
        path = alt
    return {k: (v.float() if v.is_floating_point() else v) for k, v in torch.load(path, map_location="cpu").items()}


def _c2(w, p, x, stride=1, padding=1, groups=1):
    return F.conv2d(x, w[f"{p}.weight"], w[f"{p}.bias"], stride, padding, 1, groups)


def _lin(w, p, x):
    return F.linear(x, w[f"{p}.weight"], w[f"{p}.bias"])


def encoder(w, x):
    """Encoder.forward (:222-235): layers 10..16 see [x0 | out] interleaved per group (groups 2, 4, 8, 1)."""
    bt = x.shape[0]
    out = x
    for i, (stride, groups) in zip(range(0, 18, 2), ((2, 1), (1, 1), (2, 1), (1, 1), (1, 1), (1, 2), (1, 4), (1, 8), (1, 1))):
        if i == 8:
            x0 = out
            h, wd = x0.shape[-2:]
        if i > 8:
            g = (1, 2, 4, 8, 1)[(i - 8) // 2]
            out = torch.cat((x0.view(bt, g, -1, h, wd), out.view(bt, g, -1, h, wd)), 2).view(bt, -1, h, wd)
        out = F.leaky_relu(_c2(w, f"encoder.layers.{i}", out, stride, 1, groups), 0.2)
    return out


def _deform_align(w, p, x, cond, flow):
    """DeformableAlignment.forward (:59-72): offsets 3*tanh(.) + the flow (as (dy, dx)) for each of the 16 groups x 9 taps."""
    o = cond
    for i in (0, 2, 4):
        o = F.leaky_relu(_c2(w, f"{p}.conv_offset.{i}", o), 0.1)
    o1, o2, m = torch.chunk(_c2(w, f"{p}.conv_offset.6", o), 3, 1)
    offset = 3.0 * torch.tanh(torch.cat((o1, o2), 1))
    offset = offset + flow.flip(1).repeat(1, offset.size(1) // 2, 1, 1)
    return deform_conv2d(x, offset, w[f"{p}.weight"], w[f"{p}.bias"], 1, 1, 1, torch.sigmoid(m))


def feature_propagation(w, x, flows_f, flows_b, mask, interpolation="bilinear"):
    """BidirectionalPropagation(128, learnable=True).forward (:107-193) -> the fused features [b,t,c,h,w]."""
    b, t, c, h, wd = x.shape
    p = "feat_prop_module"
    feats = [x[:, i] for i in range(t)]
    masks = [mask[:, i] for i in range(t)]
    outs = {}
    for name in ("backward_1", "forward_1"):
        order = list(range(t - 1, -1, -1)) if name == "backward_1" else list(range(t))
        res = [None] * t
        prop = None
        for n, idx in enumerate(order):
            cur, mcur = feats[idx], masks[idx]
            if n == 0:
                prop = cur
            else:
                fi = idx if name == "backward_1" else idx - 1
                fprop = (flows_f if name == "backward_1" else flows_b)[:, fi]
                fcheck = (flows_b if name == "backward_1" else flows_f)[:, fi]
                valid = P.fb_consistency(fprop, fcheck)
                warped = P.flow_warp(prop, fprop.permute(0, 2, 3, 1), interpolation)
                cond = torch.cat((cur, warped, fprop, valid, mcur), 1)
                prop = _deform_align(w, f"{p}.deform_align.{name}", prop, cond, fprop)
            y = _c2(w, f"{p}.backbone.{name}.2", F.leaky_relu(_c2(w, f"{p}.backbone.{name}.0", torch.cat((cur, prop, mcur), 1)), 0.2))
            prop = prop + y
            res[idx] = prop
        outs[name] = res
        feats = res                      # the forward pass runs over the backward pass's output
    ob = torch.stack(outs["backward_1"], 1).view(-1, c, h, wd)
    of = torch.stack(outs["forward_1"], 1).view(-1, c, h, wd)
    z = torch.cat((ob, of, mask.reshape(-1, 2, h, wd)), 1)
    y = _c2(w, f"{p}.fuse.2", F.leaky_relu(_c2(w, f"{p}.fuse.0", z), 0.2)) + x.reshape(-1, c, h, wd)
    return y.view(b, t, c, h, wd)


def _partition(x, nh):
    B, T, H, W, C = x.shape
    wh, ww = WINDOW
    x = x.view(B, T, H // wh, wh, W // ww, ww, nh, C // nh).permute(0, 2, 4, 6, 1, 3, 5, 7).contiguous()
    return x.view(B, (H // wh) * (W // ww), nh, T, wh * ww, C // nh)


def window_attention(w, p, x, mask, t_ind):
    """SparseWindowAttention.forward (:168-283); x [b,t,h,w,c] tokens, mask [b,l_t,h,w,1]."""
    b, t, h, wd, c = x.shape
    wh, ww = WINDOW
    ch = c // HEADS
    nwh, nww = math.ceil(h / wh), math.ceil(wd / ww)
    nh_, nw_ = nwh * wh, nww * ww
    if nh_ > h or nw_ > wd:
        x = F.pad(x, (0, 0, 0, nw_ - wd, 0, nh_ - h))
        mask = F.pad(mask, (0, 0, 0, nw_ - wd, 0, nh_ - h))
    q, k, v = _lin(w, f"{p}.query", x), _lin(w, f"{p}.key", x), _lin(w, f"{p}.value", x)
    wq, wk, wv = _partition(q, HEADS), _partition(k, HEADS), _partition(v, HEADS)
    eh, ew = (wh + 1) // 2, (ww + 1) // 2
    shifts = ((-eh, -ew), (-eh, ew), (eh, -ew), (eh, ew))
    idx = w[f"{p}.valid_ind_rolled"].long()
    rk = torch.cat([_partition(torch.roll(k, s, (2, 3)), HEADS) for s in shifts], 4)[:, :, :, :, idx]
    rv = torch.cat([_partition(torch.roll(v, s, (2, 3)), HEADS) for s in shifts], 4)[:, :, :, :, idx]
    pooled = F.conv2d(x.view(b * t, nh_, nw_, c).permute(0, 3, 1, 2), w[f"{p}.pool_layer.weight"], w[f"{p}.pool_layer.bias"], POOL, 0, 1, c)
    ph, pw = pooled.shape[-2:]
    pooled = pooled.permute(0, 2, 3, 1).view(b, t, ph, pw, c)

    def pool_tokens(name):
        z = _lin(w, f"{p}.{name}", pooled).view(b, 1, t, ph * pw, HEADS, ch).permute(0, 1, 4, 2, 3, 5)
        return z.expand(b, nwh * nww, HEADS, t, ph * pw, ch)

    wk = torch.cat((wk, rk, pool_tokens("key")), 4)
    wv = torch.cat((wv, rv, pool_tokens("value")), 4)
    out = torch.zeros_like(wq)
    lt = mask.shape[1]
    m = F.max_pool2d(mask.reshape(b * lt, 1, nh_, nw_), WINDOW, WINDOW).view(b, lt, nwh * nww).sum(1)
    scale = 1.0 / math.sqrt(ch)
    for i in range(b):
        mi = m[i].nonzero(as_tuple=False).view(-1)
        if len(mi):
            qt = wq[i, mi].reshape(len(mi), HEADS, t * wh * ww, ch)
            kt = wk[i, mi][:, :, t_ind].reshape(len(mi), HEADS, -1, ch)
            vt = wv[i, mi][:, :, t_ind].reshape(len(mi), HEADS, -1, ch)
            att = F.softmax((qt @ kt.transpose(-2, -1)) * scale, -1)
            out[i, mi] = (att @ vt).view(-1, HEADS, t, wh * ww, ch)
        ui = (m[i] == 0).nonzero(as_tuple=False).view(-1)
        qs, ks, vs = wq[i, ui], wk[i, ui, :, :, :wh * ww], wv[i, ui, :, :, :wh * ww]
        out[i, ui] = F.softmax((qs @ ks.transpose(-2, -1)) * scale, -1) @ vs
    out = out.view(b, nwh, nww, HEADS, t, wh, ww, ch).permute(0, 4, 1, 5, 2, 6, 3, 7).contiguous().view(b, t, nh_, nw_, c)
    return _lin(w, f"{p}.proj", out[:, :, :h, :wd])


def fusion_ffn(w, p, x, out_size):
    """FusionFeedForward.forward (:84-112): fc1 -> fold / overlap-normalise / unfold -> GELU -> fc2."""
    b, n, _ = x.shape
    nv = 1
    for i, d in enumerate(T2T["kernel_size"]):
        nv *= int((out_size[i] + 2 * T2T["padding"][i] - (d - 1) - 1) / T2T["stride"][i] + 1)
    x = _lin(w, f"{p}.fc1.0", x)
    c = x.shape[2]
    norm = F.fold(x.new_ones(b, n, 49).view(-1, nv, 49).permute(0, 2, 1), out_size, **T2T)
    x = F.fold(x.view(-1, nv, c).permute(0, 2, 1), out_size, **T2T)
    x = F.unfold(x / norm, **T2T).permute(0, 2, 1).contiguous().view(b, n, c)
    return _lin(w, f"{p}.fc2.1", F.gelu(x))


def transformer(w, x, fold_size, mask):
    """TemporalSparseTransformerBlock.forward (:327-344), t_dilation = 2."""
    B, T, H, W, C = x.shape
    t_inds = [torch.arange(i, T, 2) for i in range(2)] * (DEPTH // 2)
    for i in range(DEPTH):
        p = f"transformers.transformer.{i}"
        y = F.layer_norm(x, (C,), w[f"{p}.norm1.weight"], w[f"{p}.norm1.bias"])
        x = x + window_attention(w, f"{p}.attention", y, mask, t_inds[i])
        y = F.layer_norm(x, (C,), w[f"{p}.norm2.weight"], w[f"{p}.norm2.bias"])
        x = x + fusion_ffn(w, f"{p}.mlp", y.view(B, T * H * W, C), fold_size).view(B, T, H, W, C)
    return x


def _up(w, p, x):
    return _c2(w, f"{p}.conv", F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True))


def generator(w, masked_frames, flows_f, flows_b, masks_in, masks_updated, l_t, taps: dict | None = None):
    """InpaintGenerator.forward in eval mode -> [b, l_t, 3, H, W] in [-1, 1].  `taps` receives stage outputs (device bring-up)."""
    def tap(name, t):
        if taps is not None:
            taps[name] = t.clone()

    with torch.no_grad():
        b, t, _, H, W = masked_frames.shape
        enc = encoder(w, torch.cat((masked_frames.view(b * t, 3, H, W), masks_in.view(b * t, 1, H, W), masks_updated.view(b * t, 1, H, W)), 1))
        c, h, wd = enc.shape[1:]
        enc = enc.view(b, t, c, h, wd)
        tap("enc", enc)
        ds_f = F.interpolate(flows_f.reshape(-1, 2, H, W), scale_factor=1 / 4, mode="bilinear", align_corners=False).view(b, l_t - 1, 2, h, wd) / 4.0
        ds_b = F.interpolate(flows_b.reshape(-1, 2, H, W), scale_factor=1 / 4, mode="bilinear", align_corners=False).view(b, l_t - 1, 2, h, wd) / 4.0
        ds_in = F.interpolate(masks_in.reshape(-1, 1, H, W), scale_factor=1 / 4, mode="nearest").view(b, t, 1, h, wd)
        ds_up = F.interpolate(masks_updated[:, :l_t].reshape(-1, 1, H, W), scale_factor=1 / 4, mode="nearest").view(b, l_t, 1, h, wd)
        pool = F.max_pool2d(ds_in[:, :l_t].reshape(-1, 1, h, wd), T2T["kernel_size"], T2T["stride"], T2T["padding"])
        pool = pool.view(b, l_t, 1, pool.shape[-2], pool.shape[-1]).permute(0, 1, 3, 4, 2).contiguous()
        local = feature_propagation(w, enc[:, :l_t], ds_f, ds_b, torch.cat((ds_in[:, :l_t], ds_up), 2))
        enc = torch.cat((local, enc[:, l_t:]), 1)
        tap("ds_flows_f", ds_f)
        tap("pool_mask", pool)
        tap("enc_prop", enc)
        tok = F.unfold(enc.view(-1, c, h, wd), **T2T).permute(0, 2, 1)
        tok = _lin(w, "ss.embedding", tok).view(b, -1, pool.shape[2], pool.shape[3], 512)
        tap("tokens_in", tok)
        tok = transformer(w, tok, (h, wd), pool)
        tap("tokens_out", tok)
        z = _lin(w, "sc.embedding", tok.view(b, -1, 512))
        z = F.fold(z.view(b * t, -1, z.shape[-1]).permute(0, 2, 1), (h, wd), **T2T)
        enc = enc + _c2(w, "sc.bias_conv", z).view(b, t, c, h, wd)
        tap("enc_trans", enc)
        y = enc[:, :l_t].reshape(-1, c, h, wd)
        y = F.leaky_relu(_up(w, "decoder.0", y), 0.2)
        y = F.leaky_relu(_c2(w, "decoder.2", y), 0.2)
        y = F.leaky_relu(_up(w, "decoder.4", y), 0.2)
        return torch.tanh(_c2(w, "decoder.6", y)).view(b, l_t, 3, H, W)


def sub_ranges(length: int, sub: int, pad: int) -> List[tuple]:
    """The overlapped chunks of :254-272 / :284-304: for f = 0, sub, 2 sub, ...: run on [s, e) = [f - pad, f + sub + pad) clipped to the
    sequence and keep the part [keep_s, keep_e) of the chunk's result that belongs to [f, f + sub).  -> [(s, e, keep_s, keep_e)]"""
    out = []
    for f in range(0, length, sub):
        s, e = max(0, f - pad), min(length, f + sub + pad)
        out.append((s, e, f - s, (e - s) - (e - min(length, f + sub))))
    return out


def inpaint(weights: Dict[str, Dict[str, torch.Tensor]], frames_bgr: Sequence[np.ndarray], mask: np.ndarray, raft_iters: int = P.RAFT_ITERS,
            taps: dict | None = None, sub_video_length: int = 80) -> List[np.ndarray]:
    """PropainterInpaint.inpaint (propainter_inpaint.py:192-361) with the RAFT clip lengths of :209-236 and, for sequences longer than
    `sub_video_length`, the overlapped chunks of flow completion (:251-276, pad 5) and image propagation (:281-312, chunks of
    min(100, sub_video_length) frames, pad 10) and the capped reference frames of the window loop (:321-324);
    `weights` = {"raft": ..., "rfc": ..., "gen": ...}.  Returns BGR uint8 frames."""
    from oracle import raft_oracle as R
    from oracle import rfc_oracle as C

    T = len(frames_bgr)
    H, W = frames_bgr[0].shape[:2]
    rgb = [f[:, :, ::-1] for f in frames_bgr]
    fm, md = P.read_mask(mask, T)
    x = torch.from_numpy(np.stack(rgb).astype(np.float32) / 255).permute(0, 3, 1, 2)[None] * 2 - 1
    flow_masks = torch.from_numpy(np.stack(fm).astype(np.float32) / 255)[None, :, None]
    masks = torch.from_numpy(np.stack(md).astype(np.float32) / 255)[None, :, None]
    clip = 12 if W <= 640 else 8 if W <= 720 else 4 if W <= 1280 else 2
    if T > clip:
        ff, fb = [], []
        for f in range(0, T, clip):
            a, b_ = R.raft_bi(weights["raft"], x[:, max(f - 1, 0):min(T, f + clip)], raft_iters)
            ff.append(a)
            fb.append(b_)
        gf, gb = torch.cat(ff, 1), torch.cat(fb, 1)
    else:
        gf, gb = R.raft_bi(weights["raft"], x, raft_iters)
    if gf.shape[1] > sub_video_length:
        parts = []
        for s, e, ks, ke in sub_ranges(gf.shape[1], sub_video_length, 5):
            a, b_ = C.complete_bidirectional(weights["rfc"], gf[:, s:e], gb[:, s:e], flow_masks[:, s:e + 1])
            parts.append((a[:, ks:ke], b_[:, ks:ke]))
        pf, pb = torch.cat([a for a, _ in parts], 1), torch.cat([b_ for _, b_ in parts], 1)
    else:
        pf, pb = C.complete_bidirectional(weights["rfc"], gf, gb, flow_masks)
    sub_prop = min(100, sub_video_length)
    if T > sub_prop:
        ups, ums = [], []
        for s, e, ks, ke in sub_ranges(T, sub_prop, 10):
            pr, um = P.img_propagation((x * (1 - masks))[:, s:e], pf[:, s:e - 1], pb[:, s:e - 1], masks[:, s:e])
            ups.append(P.updated_frames(x[:, s:e], masks[:, s:e], pr)[:, ks:ke])
            ums.append(um[:, ks:ke])
        updated, upd = torch.cat(ups, 1), torch.cat(ums, 1)
        prop = None
    else:
        prop, upd = P.img_propagation(x * (1 - masks), pf, pb, masks)
        updated = P.updated_frames(x, masks, prop)
    if taps is not None:
        taps.update(gt_flows_f=gf, gt_flows_b=gb, pred_flows_f=pf, pred_flows_b=pb, prop_frames=prop, prop_masks=upd)
    comp: List = [None] * T
    binary = masks[0].permute(0, 2, 3, 1).numpy().astype(np.uint8)
    for nb, refs in P.window_schedule(T, sub_video_length):
        ids = nb + refs
        pred = generator(weights["gen"], updated[:, ids], pf[:, nb[:-1]], pb[:, nb[:-1]], masks[:, ids], upd[:, ids], len(nb))
        pred = ((pred.view(-1, 3, H, W) + 1) / 2).permute(0, 2, 3, 1).numpy() * 255
        P.composite(comp, pred, binary[nb], [np.ascontiguousarray(r) for r in rgb], nb)
    return [c[:, :, ::-1].copy() for c in comp]


def propainter_call(weights, input_frames: Sequence[np.ndarray], input_mask: np.ndarray) -> List[np.ndarray]:
    """PropainterInpaint.__call__ (:363-418): strips (multiple of 8 rows), first-frame mask, every strip replaced whole."""
    mask = input_mask if input_mask.ndim == 2 else input_mask[:, :, 0]
    H, W = mask.shape
    out = [f.copy() for f in input_frames]
    for (y0, y1, x0, x1) in P.strip_areas(W, H, mask):
        comps = inpaint(weights, [f[y0:y1, x0:x1] for f in out], mask[y0:y1, x0:x1])
        for f, c in zip(out, comps):
            f[y0:y1, x0:x1] = c
    return out
