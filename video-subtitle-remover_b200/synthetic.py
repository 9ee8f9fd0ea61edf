"""Synthetic inputs of the benchmark configurations (BASELINE.json: "synthetic N-frame clip, fixed subtitle bbox"): a moving textured
background and the mask of the default selection area.  Deterministic in their arguments, so every rank and the CPU baseline leg of
bench.py see the same pixels.  (tests/ checks that these equal the generators the oracle's goldens were made with.)"""
from typing import List

import numpy as np


def synthetic_clip(n: int, H: int, W: int, seed: int = 0, pad: int = 64) -> List[np.ndarray]:
    """n BGR uint8 frames [H,W,3]: a smooth random texture (x4 bilinear up-sampling of noise) translating by (3, 2) px per frame."""
    rng = np.random.default_rng(seed)
    bh, bw = H + 2 * pad, W + 3 * pad
    coarse = rng.integers(0, 256, (bh // 4 + 2, bw // 4 + 2, 3), dtype=np.uint8).astype(np.float32)
    yy = np.linspace(0, coarse.shape[0] - 1.001, bh).astype(np.float32)
    xx = np.linspace(0, coarse.shape[1] - 1.001, bw).astype(np.float32)
    y0, x0 = np.floor(yy).astype(np.int64), np.floor(xx).astype(np.int64)
    fy, fx = (yy - y0)[:, None, None], (xx - x0)[None, :, None]
    rows0, rows1 = coarse[y0], coarse[y0 + 1]
    big = rows0[:, x0] * (1 - fy) * (1 - fx) + rows0[:, x0 + 1] * (1 - fy) * fx + rows1[:, x0] * fy * (1 - fx) + rows1[:, x0 + 1] * fy * fx
    big = np.clip(big, 0, 255).astype(np.uint8)
    return [np.ascontiguousarray(big[(2 * i) % (2 * pad):(2 * i) % (2 * pad) + H, (3 * i) % (3 * pad):(3 * i) % (3 * pad) + W]) for i in range(n)]


def default_mask(H: int, W: int, area=(0.88, 0.99, 0.15, 0.85)) -> np.ndarray:
    """Mask of the default selection area of backend/config.py:43 ("ymin,ymax,xmin,xmax" as fractions) through `create_mask`."""
    from .inpaint_tools import create_mask

    return create_mask((H, W), [(int(W * area[2]), int(W * area[3]), int(H * area[0]), int(H * area[1]))])


def sttn_weight_shapes():
    """Tensor inventory of the sttn-auto / sttn-det generator checkpoint `ckpt['netG']` (auto_sttn.py:64-95,148-222; the two networks
    differ in patch geometry, not in tensors)."""
    shapes = {}

    def conv(name, co, ci, k):
        shapes[name + ".weight"] = (co, ci, k, k)
        shapes[name + ".bias"] = (co,)

    conv("encoder.0", 64, 3, 3)
    conv("encoder.2", 64, 64, 3)
    conv("encoder.4", 128, 64, 3)
    conv("encoder.6", 256, 128, 3)
    for b in range(8):
        p = f"transformer.{b}."
        conv(p + "attention.query_embedding", 256, 256, 1)
        conv(p + "attention.value_embedding", 256, 256, 1)
        conv(p + "attention.key_embedding", 256, 256, 1)
        conv(p + "attention.output_linear.0", 256, 256, 3)
        conv(p + "feed_forward.conv.0", 256, 256, 3)
        conv(p + "feed_forward.conv.2", 256, 256, 3)
    conv("decoder.0.conv", 128, 256, 3)
    conv("decoder.2", 64, 128, 3)
    conv("decoder.4.conv", 64, 64, 3)
    conv("decoder.6", 3, 64, 3)
    return shapes


def random_sttn_weights(seed: int = 0):
    """Seeded random-init weights of that architecture (BASELINE's "random-init weights of that architecture" when no checkpoint is
    staged): gain 1/sqrt(fan_in), biases 0.05 sigma — decoded images use the full 0..255 range while features stay small."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in sttn_weight_shapes().items():
        if name.endswith(".weight"):
            out[name] = rng.standard_normal(shape, dtype=np.float32) * np.float32(1.0 / np.sqrt(shape[1] * shape[2] * shape[3]))
        else:
            out[name] = rng.standard_normal(shape, dtype=np.float32) * np.float32(0.05)
    return out
