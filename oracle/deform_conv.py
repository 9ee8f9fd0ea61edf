"""Pure-torch modulated deformable convolution — TEST INFRASTRUCTURE ONLY (oracle of the ProPainter rows P4/P6).

The reference calls `torchvision.ops.deform_conv2d(x, offset, weight, bias, stride, padding, dilation, mask)`
(backend/inpaint/video/model/recurrent_flow_completion.py:44-46, propainter.py:69-72); torchvision's CPU kernel crashes
(SIGSEGV) on the reference's shapes in this image (torchvision 0.26 / torch 2.11), so the oracle and the golden generator
(tools/make_golden_propainter.py) use this restatement of the published operator:

  out[b,co,i,j] = bias[co] + sum_{ci,k} w[co,ci,k] * m[b,g(ci)*K+k,i,j] * bilinear(x[b,ci], i*s - p + ky*d + dy, j*s - p + kx*d + dx)

with (dy, dx) = offset[b, 2*(g*K+k) + {0,1}, i, j], g(ci) the offset group of channel ci, and bilinear sampling that treats
everything outside the image as zero.  tests/test_propainter_oracle.py checks it against torchvision on shapes where
torchvision's kernel survives.
"""
import torch
import torch.nn.functional as F


def deform_conv2d(x, offset, weight, bias=None, stride=1, padding=1, dilation=1, mask=None):
    stride = (stride, stride) if isinstance(stride, int) else tuple(stride)
    padding = (padding, padding) if isinstance(padding, int) else tuple(padding)
    dilation = (dilation, dilation) if isinstance(dilation, int) else tuple(dilation)
    B, Cin, H, W = x.shape
    Cout, Cin_g, kh, kw = weight.shape
    K = kh * kw
    groups = Cin // Cin_g
    G = offset.shape[1] // (2 * K)                       # offset groups
    Ho, Wo = offset.shape[2], offset.shape[3]
    cpg = Cin // G
    dev, dt = x.device, x.dtype
    base_y = (torch.arange(Ho, device=dev, dtype=dt) * stride[0] - padding[0]).view(1, Ho, 1)
    base_x = (torch.arange(Wo, device=dev, dtype=dt) * stride[1] - padding[1]).view(1, 1, Wo)
    off = offset.view(B, G, K, 2, Ho, Wo)
    msk = mask.view(B, G, K, Ho, Wo) if mask is not None else None
    cols = x.new_empty(B, G, cpg, K, Ho, Wo)
    for g in range(G):
        xg = x[:, g * cpg:(g + 1) * cpg]
        for k in range(K):
            ky, kx = divmod(k, kw)
            py = base_y + ky * dilation[0] + off[:, g, k, 0]
            px = base_x + kx * dilation[1] + off[:, g, k, 1]
            grid = torch.stack((2 * px / max(W - 1, 1) - 1, 2 * py / max(H - 1, 1) - 1), -1)   # align_corners=True pixel coordinates
            v = F.grid_sample(xg, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
            cols[:, g, :, k] = v * msk[:, g, k].unsqueeze(1) if msk is not None else v
    cols = cols.view(B, groups, Cin_g, K, Ho * Wo)       # channel ci = g*cpg + c = group*Cin_g + c'
    w = weight.view(groups, Cout // groups, Cin_g, K)
    out = torch.einsum("bgckp,gock->bgop", cols, w).reshape(B, Cout, Ho, Wo)
    if bias is not None:
        out = out + bias.view(1, -1, 1, 1)
    return out
