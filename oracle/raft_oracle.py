"""CPU oracle for RAFT as ProPainter uses it — TEST INFRASTRUCTURE ONLY.

Restates, functionally over the `raft-things.pth` state dict (keys `module.*`, DataParallel prefix):
  RAFT_bi.forward                  backend/inpaint/video/model/modules/flow_comp_raft.py:27-55  (frame t -> t+1 and t+1 -> t)
  RAFT.forward (large model)       backend/inpaint/video/raft/raft.py:87-146
  BasicEncoder / ResidualBlock     raft/extractor.py:6-58,118-190   (fnet: instance norm, cnet: batch norm)
  CorrBlock                        raft/corr.py:12-60               (all-pairs / sqrt(256), 3 x avg-pool, 9x9 lookups x 4 levels)
  BasicUpdateBlock                 raft/update.py:60-139            (motion encoder, SepConvGRU, flow head, 0.25 * mask head)
  upsample_flow                    raft/raft.py:73-84               (convex 8x up-sampling)
Parity: PINNED against tests/golden/propainter_real.npz (`gt_flows_f/b`: RAFT_bi of the unmodified reference, 20 iterations).
Note (reference behaviour, kept): the lookup window adds the meshgrid's dy component to x and dx to y (corr.py:37-43), and
inputs below 128 rows make level 3 of the pyramid one row high, which turns its sample coordinates into 0/0.
"""
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

CORR_LEVELS, CORR_RADIUS, HDIM, CDIM = 4, 4, 128, 128     # raft.py:37-41 (args.small = False)


def load_weights(path: str) -> Dict[str, torch.Tensor]:
    sd = torch.load(path, map_location="cpu")
    return {k[len("module."):] if k.startswith("module.") else k: v.float() for k, v in sd.items()}


def _conv(w, p, x, stride=1, padding=0):
    return F.conv2d(x, w[f"{p}.weight"], w.get(f"{p}.bias"), stride, padding)


def _norm(w, p, x, kind):
    if kind == "instance":          # nn.InstanceNorm2d default: no affine, no running statistics
        return F.instance_norm(x)
    return F.batch_norm(x, w[f"{p}.running_mean"], w[f"{p}.running_var"], w[f"{p}.weight"], w[f"{p}.bias"], False, 0.0, 1e-5)


def _res_block(w, p, x, kind, stride):
    y = F.relu(_norm(w, f"{p}.norm1", _conv(w, f"{p}.conv1", x, stride, 1), kind))
    y = F.relu(_norm(w, f"{p}.norm2", _conv(w, f"{p}.conv2", y, 1, 1), kind))
    if stride != 1:                 # downsample = Sequential(conv1x1 stride, norm3); norm3 is also registered as `{p}.norm3`
        x = _norm(w, f"{p}.downsample.1", _conv(w, f"{p}.downsample.0", x, stride), kind)
    return F.relu(x + y)


def encoder(w, p, x, kind):
    """BasicEncoder.forward: /8 resolution features."""
    x = F.relu(_norm(w, f"{p}.norm1", _conv(w, f"{p}.conv1", x, 2, 3), kind))
    for layer, stride in (("layer1", 1), ("layer2", 2), ("layer3", 2)):
        x = _res_block(w, f"{p}.{layer}.0", x, kind, stride)
        x = _res_block(w, f"{p}.{layer}.1", x, kind, 1)
    return _conv(w, f"{p}.conv2", x)


def coords_grid(n, h, w_):
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w_), indexing="ij")
    return torch.stack((xs, ys), 0).float()[None].repeat(n, 1, 1, 1)


def corr_pyramid(f1, f2):
    n, d, h, w_ = f1.shape
    c = torch.matmul(f1.view(n, d, h * w_).transpose(1, 2), f2.view(n, d, h * w_)) / torch.sqrt(torch.tensor(d).float())
    c = c.view(n * h * w_, 1, h, w_)
    pyr = [c]
    for _ in range(CORR_LEVELS - 1):
        c = F.avg_pool2d(c, 2, stride=2)
        pyr.append(c)
    return pyr


def corr_lookup(pyr, coords):
    r = CORR_RADIUS
    n, _, h, w_ = coords.shape
    c = coords.permute(0, 2, 3, 1).reshape(n * h * w_, 1, 1, 2)
    d = torch.linspace(-r, r, 2 * r + 1)
    dy, dx = torch.meshgrid(d, d, indexing="ij")
    delta = torch.stack((dy, dx), -1).view(1, 2 * r + 1, 2 * r + 1, 2)      # (dy, dx) added to (x, y): the reference's order
    out = []
    for i, lvl in enumerate(pyr):
        p = c / 2 ** i + delta
        H, W = lvl.shape[-2:]
        grid = torch.cat((2 * p[..., :1] / (W - 1) - 1, 2 * p[..., 1:] / (H - 1) - 1), -1)
        out.append(F.grid_sample(lvl, grid, align_corners=True).view(n, h, w_, -1))
    return torch.cat(out, -1).permute(0, 3, 1, 2).contiguous()


def update_block(w, net, inp, corr, flow):
    p = "update_block"
    cor = F.relu(_conv(w, f"{p}.encoder.convc1", corr))
    cor = F.relu(_conv(w, f"{p}.encoder.convc2", cor, 1, 1))
    flo = F.relu(_conv(w, f"{p}.encoder.convf1", flow, 1, 3))
    flo = F.relu(_conv(w, f"{p}.encoder.convf2", flo, 1, 1))
    mot = torch.cat((F.relu(_conv(w, f"{p}.encoder.conv", torch.cat((cor, flo), 1), 1, 1)), flow), 1)
    x = torch.cat((inp, mot), 1)
    h = net
    for s, pad in (("1", (0, 2)), ("2", (2, 0))):          # SepConvGRU: horizontal (1x5) then vertical (5x1)
        hx = torch.cat((h, x), 1)
        z = torch.sigmoid(_conv(w, f"{p}.gru.convz{s}", hx, 1, pad))
        r = torch.sigmoid(_conv(w, f"{p}.gru.convr{s}", hx, 1, pad))
        q = torch.tanh(_conv(w, f"{p}.gru.convq{s}", torch.cat((r * h, x), 1), 1, pad))
        h = (1 - z) * h + z * q
    delta = _conv(w, f"{p}.flow_head.conv2", F.relu(_conv(w, f"{p}.flow_head.conv1", h, 1, 1)), 1, 1)
    mask = 0.25 * _conv(w, f"{p}.mask.2", F.relu(_conv(w, f"{p}.mask.0", h, 1, 1)))
    return h, mask, delta


def upsample_flow(flow, mask):
    n, _, h, w_ = flow.shape
    m = torch.softmax(mask.view(n, 1, 9, 8, 8, h, w_), 2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(n, 2, 9, 1, 1, h, w_)
    return torch.sum(m * up, 2).permute(0, 1, 4, 2, 5, 3).reshape(n, 2, 8 * h, 8 * w_)


def raft(w, image1, image2, iters=20) -> torch.Tensor:
    """RAFT.forward(test_mode=True)[1]: full-resolution flow image1 -> image2; images [N,3,H,W] in [-1, 1]."""
    with torch.no_grad():
        f = encoder(w, "fnet", torch.cat((image1, image2), 0), "instance")
        f1, f2 = torch.split(f, [image1.shape[0]] * 2, 0)
        pyr = corr_pyramid(f1.float(), f2.float())
        c = encoder(w, "cnet", image1, "batch")
        net, inp = torch.tanh(c[:, :HDIM]), torch.relu(c[:, HDIM:HDIM + CDIM])
        n, _, H, W = image1.shape
        coords0 = coords_grid(n, H // 8, W // 8)
        coords1 = coords0.clone()
        up = None
        for _ in range(iters):
            corr = corr_lookup(pyr, coords1)
            net, mask, delta = update_block(w, net, inp, corr, coords1 - coords0)
            coords1 = coords1 + delta
            up = upsample_flow(coords1 - coords0, mask)
        return up


def raft_bi(w, frames: torch.Tensor, iters=20) -> Tuple[torch.Tensor, torch.Tensor]:
    """RAFT_bi.forward: frames [b,t,3,h,w] -> (forward flows t -> t+1, backward flows t+1 -> t), each [b,t-1,2,h,w]."""
    b, t, c, h, w_ = frames.shape
    a, bb = frames[:, :-1].reshape(-1, c, h, w_), frames[:, 1:].reshape(-1, c, h, w_)
    return raft(w, a, bb, iters).view(b, t - 1, 2, h, w_), raft(w, bb, a, iters).view(b, t - 1, 2, h, w_)
