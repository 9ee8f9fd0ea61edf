"""CPU oracle for ProPainter's flow completion — TEST INFRASTRUCTURE ONLY.

Restates, functionally over `recurrent_flow_completion.pth`:
  RecurrentFlowCompleteNet.forward / forward_bidirect_flow / combine_flow   video/model/recurrent_flow_completion.py:275-348
  downsample (Conv3d 1x5x5 s2, replicate padding), P3DBlock encoders, dilated mid convs                     :150-173,212-240
  BidirectionalPropagation with SecondOrderDeformableAlignment (16 deformable groups, offsets 5*tanh)       :10-128
  decoders (`deconv` = bilinear x2 align_corners + conv3x3)                                                 :130-148,245-262
The deformable convolution is oracle/deform_conv.py.  The edge head only runs in training mode and is not restated.
Parity: PINNED against tests/golden/propainter_real.npz (`pred_flows_f/b`, produced by the unmodified reference from `gt_flows_*`).
"""
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from oracle.deform_conv import deform_conv2d

MAX_RESIDUE = 5.0      # SecondOrderDeformableAlignment.max_residue_magnitude (:14)
DEFORM_GROUPS = 16     # :58


def load_weights(path: str) -> Dict[str, torch.Tensor]:
    return {k: v.float() for k, v in torch.load(path, map_location="cpu").items()}


def _c2(w, p, x, stride=1, padding=1, dilation=1):
    return F.conv2d(x, w[f"{p}.weight"], w[f"{p}.bias"], stride, padding, dilation)


def _c3(w, p, x, stride=(1, 1, 1), padding=(0, 0, 0), dilation=(1, 1, 1)):
    return F.conv3d(x, w[f"{p}.weight"], w[f"{p}.bias"], stride, padding, dilation)


def _p3d(w, p, x, stride):
    """P3DBlock (use_residual=0): spatial 1x3x3 conv + LeakyReLU(0.2), then temporal 3x1x1 conv with dilation 2."""
    y = F.leaky_relu(_c3(w, f"{p}.conv1.0", x, (1, stride, stride), (0, 1, 1)), 0.2)
    return _c3(w, f"{p}.conv2.0", y, (1, 1, 1), (2, 0, 0), (2, 1, 1))


def _deform_align(w, p, x, cond):
    """SecondOrderDeformableAlignment.forward (:33-46): x = [feat(t-1), feat(t-2)] (256 ch), cond = 384 ch."""
    o = cond
    for i in (0, 2, 4):
        o = F.leaky_relu(_c2(w, f"{p}.conv_offset.{i}", o), 0.1)
    o = _c2(w, f"{p}.conv_offset.6", o)
    o1, o2, m = torch.chunk(o, 3, 1)
    offset = MAX_RESIDUE * torch.tanh(torch.cat((o1, o2), 1))
    return deform_conv2d(x, offset, w[f"{p}.weight"], w[f"{p}.bias"], 1, 1, 1, torch.sigmoid(m))


def _propagate(w, x):
    """BidirectionalPropagation.forward (:70-128): x [b,t,c,h,w]."""
    b, t, c, h, wd = x.shape
    p = "feat_prop_module"
    spatial = [x[:, i] for i in range(t)]
    feats = {"spatial": spatial}
    for name in ("backward_", "forward_"):
        feats[name] = []
        order = list(range(t - 1, -1, -1)) if name == "backward_" else list(range(t))
        prop = x.new_zeros(b, c, h, wd)
        for i, idx in enumerate(order):
            cur = spatial[idx]
            if i > 0:
                n2 = feats[name][-2] if i > 1 else torch.zeros_like(prop)
                cond = torch.cat((prop, cur, n2), 1)
                prop = _deform_align(w, f"{p}.deform_align.{name}", torch.cat((prop, n2), 1), cond)
            # [current] + [features of the directions already finished, time-ordered] + [propagated]
            feat = [cur] + [feats[k][idx] for k in feats if k not in ("spatial", name)] + [prop]
            y = _c2(w, f"{p}.backbone.{name}.2", F.leaky_relu(_c2(w, f"{p}.backbone.{name}.0", torch.cat(feat, 1)), 0.1))
            prop = prop + y
            feats[name].append(prop)
        if name == "backward_":
            feats[name] = feats[name][::-1]
    out = [F.conv2d(torch.cat((feats["backward_"][i], feats["forward_"][i]), 1), w[f"{p}.fusion.weight"], w[f"{p}.fusion.bias"]) for i in range(t)]
    return torch.stack(out, 1) + x


def _up(w, p, x):
    return _c2(w, f"{p}.conv", F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True))


def complete(w, masked_flows, masks):
    """RecurrentFlowCompleteNet.forward in eval mode: masked_flows [b,t,2,h,w], masks [b,t,1,h,w] -> flows [b,t,2,h,w]."""
    b, t, _, h, wd = masked_flows.shape
    x = torch.cat((masked_flows.permute(0, 2, 1, 3, 4), masks.permute(0, 2, 1, 3, 4)), 1)
    x = F.pad(x, (2, 2, 2, 2, 0, 0), mode="replicate")                      # Conv3d(padding_mode='replicate', padding=(0,2,2))
    x = F.leaky_relu(_c3(w, "downsample.0", x, (1, 2, 2)), 0.2)
    e1 = F.leaky_relu(_p3d(w, "encoder1.2", F.leaky_relu(_p3d(w, "encoder1.0", x, 1), 0.2), 2), 0.2)
    e2 = F.leaky_relu(_p3d(w, "encoder2.2", F.leaky_relu(_p3d(w, "encoder2.0", e1, 1), 0.2), 2), 0.2)
    m = e2
    for i, d in ((0, 3), (2, 2), (4, 1)):
        m = F.leaky_relu(_c3(w, f"mid_dilation.{i}", m, (1, 1, 1), (0, d, d), (1, d, d)), 0.2)
    f = _propagate(w, m.permute(0, 2, 1, 3, 4)).reshape(-1, 128, h // 8, wd // 8)
    e1 = e1.permute(0, 2, 1, 3, 4).reshape(-1, 64, h // 4, wd // 4)
    d2 = F.leaky_relu(_up(w, "decoder2.2", F.leaky_relu(_c2(w, "decoder2.0", f), 0.2)), 0.2) + e1
    d1 = F.leaky_relu(_up(w, "decoder1.2", F.leaky_relu(_c2(w, "decoder1.0", d2), 0.2)), 0.2)
    flow = _up(w, "upsample.2", F.leaky_relu(_c2(w, "upsample.0", d1), 0.2))
    return flow.view(b, t, 2, h, wd)


def complete_bidirectional(w, flows_f, flows_b, masks) -> Tuple[torch.Tensor, torch.Tensor]:
    """forward_bidirect_flow + combine_flow (:302-348): flows [b,t-1,2,h,w], masks [b,t,1,h,w] (1 = hole, the dilated flow mask)."""
    with torch.no_grad():
        mf, mb = masks[:, :-1].contiguous(), masks[:, 1:].contiguous()
        pf = complete(w, flows_f * (1 - mf), mf)
        pb = torch.flip(complete(w, torch.flip(flows_b * (1 - mb), [1]), torch.flip(mb, [1])), [1])
        return pf * mf + flows_f * (1 - mf), pb * mb + flows_b * (1 - mb)
