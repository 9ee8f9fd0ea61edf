"""ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE) — CPU restatement of the reference STTN-auto path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module; the product path (vsr_b200.*) never does and fails loudly without its CUDA library.

Every function names the reference lines it follows (paths relative to /root/reference).  Parity status:
PINNED — tests/test_oracle_pinned.py runs this module against the unmodified reference (imported with
oracle/ref_import.py) with the reference's real weights in the build container, and
tests/golden/*.npz hold outputs of the unmodified reference that travel to the GPU box.

Numerics: fp32 throughout, like the reference's CPU path.  The u8 arithmetic (cv2 fixed-point resize,
quantise-by-truncation, mask composite) is restated in integer numpy and is bit-exact.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# constants of the path
# ----------------------------------------------------------------------------------------------
MODEL_W, MODEL_H = 640, 120  # backend/inpaint/sttn_auto_inpaint.py:38
PATCHSIZE = [(80, 15), (32, 6), (10, 5), (5, 3)]  # (width, height) backend/inpaint/sttn/auto_sttn.py:69
CHANNEL = 256  # auto_sttn.py:67
STACK_NUM = 8  # auto_sttn.py:68
NEIGHBOR_STRIDE = 5  # backend/config.py:89
REF_LENGTH = 10  # backend/config.py:91
MAX_LOAD_NUM = 50  # backend/config.py:93-94
AREA_DEVIATION_PIXEL = 10  # backend/config.py:61


# ----------------------------------------------------------------------------------------------
# A1 / A2 / A6: integer mask + index path (bit-exact)
# ----------------------------------------------------------------------------------------------
def create_mask(size: Tuple[int, int], coords_list, deviation: int = AREA_DEVIATION_PIXEL) -> np.ndarray:
    """backend/tools/inpaint_tools.py:31-47.  Rectangles (xmin,xmax,ymin,ymax) grown by `deviation`
    px, filled with 255.  cv2.rectangle(thickness=-1) fills the INCLUSIVE box [x1..x2]x[y1..y2]
    clipped to the image; x1/y1 are clamped at 0 by the reference, x2/y2 are not (cv2 clips)."""
    h, w = size
    mask = np.zeros((h, w), dtype=np.uint8)
    for xmin, xmax, ymin, ymax in coords_list or []:
        x1, y1 = max(0, xmin - deviation), max(0, ymin - deviation)
        x2, y2 = xmax + deviation, ymax + deviation
        # cv2.rectangle orders the corners itself, so an inverted box still fills
        xa, xb = min(x1, x2), max(x1, x2)
        ya, yb = min(y1, y2), max(y1, y2)
        xa, ya = max(xa, 0), max(ya, 0)
        xb, yb = min(xb, w - 1), min(yb, h - 1)
        if xa <= xb and ya <= yb:
            mask[ya:yb + 1, xa:xb + 1] = 255
    return mask


def _components8(binary: np.ndarray):
    """8-connected components with (top, height, area, centroid_y) per component, in the label order
    of cv2.connectedComponentsWithStats(…, connectivity=8) (inpaint_tools.py:76).  OpenCV >= 4.5
    labels with the block-based Spaghetti/BBDT scan: provisional labels are created per 2x2 block in
    block-raster order and flattened in that order, so components are numbered by their first 2x2
    block (min over pixels of (y//2, x//2)), not by their first pixel.  The order only matters for
    ties of the integer centre in the stable sort at inpaint_tools.py:99."""
    from scipy import ndimage

    labels, n = ndimage.label(binary > 0, structure=np.ones((3, 3), dtype=bool))
    out = []
    for i in range(1, n + 1):
        ys, xs = np.nonzero(labels == i)
        key = int(((ys // 2).astype(np.int64) * (binary.shape[1] // 2 + 2) + xs // 2).min())
        out.append((key, int(ys.min()), int(ys.max() - ys.min() + 1), int(ys.size), float(ys.astype(np.float64).mean())))
    out.sort(key=lambda t: t[0])
    return [t[1:] for t in out]


def get_inpaint_area_by_mask(W: int, H: int, h: int, mask: np.ndarray, multiple: int = 1) -> List[Tuple[int, int, int, int]]:
    """backend/tools/inpaint_tools.py:49-242.  mask: [H,W] or [H,W,1], non-zero = subtitle.
    Returns [(ymin, ymax, xmin, xmax)] strips of height exactly h (before the `multiple` fix-up)."""
    m = np.asarray(mask)
    if m.ndim == 3:
        m = m[:, :, 0]
    if not np.any(m):
        return []
    binary = (m > 0)
    islands = []
    for top, height, area, cy in _components8(binary):
        if area < 10:  # :88
            continue
        islands.append((top, top + height, int(cy), area))
    if not islands:
        return []
    islands.sort(key=lambda t: t[2])  # stable sort by integer centre :99
    groups, cur = [], [islands[0]]
    for isl in islands[1:]:
        lo = min(i[0] for i in cur)
        hi = max(i[1] for i in cur)
        top, bot = isl[0], isl[1]
        nlo, nhi = min(lo, top), max(hi, bot)
        if hi < top:  # :119-125 gap rows must contain mask pixels
            connected = bool(np.any(binary[hi:top, :]))
        else:
            connected = True
        if nhi - nlo <= h and connected:
            cur.append(isl)
        else:
            groups.append(cur)
            cur = [isl]
    groups.append(cur)

    areas: List[Tuple[int, int, int, int]] = []
    for g in groups:
        lo = min(i[0] for i in g)
        hi = max(i[1] for i in g)
        cy = sum(i[2] for i in g) // len(g)
        half = h // 2
        ymin = max(0, cy - half)
        ymax = ymin + h
        if ymax > H:
            ymax = H
            ymin = max(0, H - h)
        if ymin > lo or ymax < hi:  # :164-184
            if hi - lo <= h:
                ymin = lo
                ymax = ymin + h
                if ymax > H:
                    ymax = H
                    ymin = max(0, H - h)
            else:
                c = (lo + hi) // 2
                ymin = max(0, c - half)
                ymax = ymin + h
                if ymax > H:
                    ymax = H
                    ymin = max(0, H - h)
        xmin, xmax = 0, W
        if multiple > 1:  # :189-235 (ProPainter)
            height = ymax - ymin
            rem = height % multiple
            if rem != 0:
                adj = multiple - rem
                c = (ymin + ymax) / 2
                if ymin - adj / 2 >= 0 and ymax + adj / 2 <= H:
                    ymin = int(c - height / 2 - adj / 2)
                    ymax = int(c + height / 2 + adj / 2)
                elif height > multiple:
                    ymin = int(c - (height - rem) / 2)
                    ymax = int(c + (height - rem) / 2)
                else:
                    if ymax + adj <= H:
                        ymax += adj
                    elif ymin - adj >= 0:
                        ymin -= adj
                    elif height > multiple:
                        ymax = ymin + height - rem
            width = xmax - xmin
            remw = width % multiple
            if remw != 0:
                cx = (xmin + xmax) / 2
                xmin = int(cx - (width - remw) / 2)
                xmax = int(cx + (width - remw) / 2)
        area = (int(ymin), int(ymax), int(xmin), int(xmax))
        if area not in areas:
            areas.append(area)
    return areas


def batch_generator(n_samples: int, max_batch_size: int) -> List[Tuple[int, int]]:
    """backend/tools/inpaint_tools.py:7-29 on index ranges: returns [(start, stop)] slices."""
    bs = max_batch_size
    nb = n_samples // bs
    while n_samples % bs < bs / 2.0 and bs > 1:
        bs -= 1
        nb = n_samples // bs
    out = [(i * bs, (i + 1) * bs) for i in range(nb)]
    if nb * bs < n_samples:
        out.append((nb * bs, n_samples))
    return out


def window_schedule(T: int, stride: int = NEIGHBOR_STRIDE, ref_length: int = REF_LENGTH):
    """backend/inpaint/sttn_auto_inpaint.py:142-146 + get_ref_index :107-120.
    Returns [(neighbor_ids, ref_ids)] in visiting order."""
    sched = []
    for f in range(0, T, stride):
        nb = list(range(max(0, f - stride), min(T, f + stride + 1)))
        refs = [i for i in range(0, T, ref_length) if i not in nb]
        sched.append((nb, refs))
    return sched


# ----------------------------------------------------------------------------------------------
# cv2.resize restatements
# ----------------------------------------------------------------------------------------------
def _linear_coeffs(src_n: int, dst_n: int, vertical: bool = False):
    """Source taps + fractional weight per destination index, as cv::resize computes them
    (imgproc resize.cpp, INTER_LINEAR): scale = 1/(dst/src) in double, f = float((d+0.5)*scale-0.5).
    Horizontal taps clamp the *coefficient* at the borders (fx=0); vertical taps keep the fraction and
    clamp only the row *indices* — visible as two separate truncations in the u8 path when up-scaling."""
    scale = 1.0 / (dst_n / src_n)  # double
    d = np.arange(dst_n, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    i0 = np.floor(f).astype(np.int64)
    a = (f - i0.astype(np.float32)).astype(np.float32)
    if vertical:
        i1 = np.clip(i0 + 1, 0, src_n - 1)
        i0 = np.clip(i0, 0, src_n - 1)
        return i0, i1, a
    lo = i0 < 0
    i0[lo] = 0
    a[lo] = 0
    hi = i0 >= src_n - 1
    i0[hi] = src_n - 1
    a[hi] = 0
    i1 = np.minimum(i0 + 1, src_n - 1)
    return i0, i1, a


def cv2_resize_linear_u8(src: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    """cv2.resize(src_u8, (dst_w, dst_h)) INTER_LINEAR, 11-bit fixed point, as called at
    sttn_auto_inpaint.py:72,270 (down-scale of the strip) and :86,312 when a comp is still uint8."""
    assert src.dtype == np.uint8
    s = src if src.ndim == 3 else src[:, :, None]
    sh, sw = s.shape[:2]
    x0, x1, ax = _linear_coeffs(sw, dst_w)
    y0, y1, ay = _linear_coeffs(sh, dst_h, vertical=True)
    wx1 = np.rint(ax * np.float32(2048)).astype(np.int32)
    wx0 = np.rint((np.float32(1) - ax) * np.float32(2048)).astype(np.int32)
    wy1 = np.rint(ay * np.float32(2048)).astype(np.int32)
    wy0 = np.rint((np.float32(1) - ay) * np.float32(2048)).astype(np.int32)
    si = s.astype(np.int32)
    hp = si[:, x0, :] * wx0[None, :, None] + si[:, x1, :] * wx1[None, :, None]  # [sh, dw, c] int32
    r0 = hp[y0]
    r1 = hp[y1]
    out = (((wy0[:, None, None] * (r0 >> 4)) >> 16) + ((wy1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out if src.ndim == 3 else out[:, :, 0]


def cv2_resize_linear_f32(src: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    """cv2.resize(src_f32, (dst_w, dst_h)) INTER_LINEAR: half-pixel bilinear in float32
    (sttn_auto_inpaint.py:86,312 once a comp has been blended to float32)."""
    assert src.dtype == np.float32
    sh, sw = src.shape[:2]
    x0, x1, ax = _linear_coeffs(sw, dst_w)
    y0, y1, ay = _linear_coeffs(sh, dst_h, vertical=True)
    ax = ax[None, :, None]
    ay = ay[:, None, None]
    one = np.float32(1)
    hp = src[:, x0, :] * (one - ax) + src[:, x1, :] * ax
    return (hp[y0] * (one - ay) + hp[y1] * ay).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------------------------
def weight_shapes() -> Dict[str, Tuple[int, ...]]:
    """Tensor inventory of ckpt['netG'] for sttn-auto (auto_sttn.py:64-95,148-222)."""
    s: Dict[str, Tuple[int, ...]] = {}

    def conv(name, co, ci, k):
        s[name + ".weight"] = (co, ci, k, k)
        s[name + ".bias"] = (co,)

    conv("encoder.0", 64, 3, 3)
    conv("encoder.2", 64, 64, 3)
    conv("encoder.4", 128, 64, 3)
    conv("encoder.6", 256, 128, 3)
    for b in range(STACK_NUM):
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
    return s


def random_weights(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded random-init weights of the sttn-auto architecture (numpy RNG so the same tensors can be
    rebuilt anywhere).  Gain 1/sqrt(fan_in): the decoded images use the full 0..255 range (std ~44) while the
    features stay < 64, so the random network is a meaningful, if chaotic, parity target."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in weight_shapes().items():
        if name.endswith(".weight"):
            fan_in = shape[1] * shape[2] * shape[3]
            w = rng.standard_normal(shape, dtype=np.float32) * np.float32(1.0 / math.sqrt(fan_in))
        else:
            w = rng.standard_normal(shape, dtype=np.float32) * np.float32(0.05)
        out[name] = torch.from_numpy(w)
    return out


def load_weights(path: str) -> Dict[str, torch.Tensor]:
    """sttn_auto_inpaint.py:34 — torch pickle {'netG': state_dict}; also accepts the .npz written by
    tools/stage_weights.py."""
    if path.endswith(".npz"):
        z = np.load(path)
        return {k: torch.from_numpy(z[k]) for k in z.files}
    sd = torch.load(path, map_location="cpu", weights_only=False)["netG"]
    return {k: v.float() for k, v in sd.items()}


# ----------------------------------------------------------------------------------------------
# network (A5, A7-A10), fp32 NCHW
# ----------------------------------------------------------------------------------------------
def _lrelu(x):
    return F.leaky_relu(x, 0.2)


def encoder(w, x: torch.Tensor) -> torch.Tensor:
    """auto_sttn.py:75-84: conv3x3 s2,s1,s2,s1 (3->64->64->128->256), LeakyReLU(0.2) after each."""
    for name, stride in (("encoder.0", 2), ("encoder.2", 1), ("encoder.4", 2), ("encoder.6", 1)):
        x = _lrelu(F.conv2d(x, w[name + ".weight"], w[name + ".bias"], stride=stride, padding=1))
    return x


def _split_tokens(x: torch.Tensor, pw: int, ph: int) -> torch.Tensor:
    """[t, d, H, W] -> [t*oh*ow, d*ph*pw] tokens (auto_sttn.py:182-190)."""
    t, d, H, W = x.shape
    oh, ow = H // ph, W // pw
    return x.reshape(t, d, oh, ph, ow, pw).permute(0, 2, 4, 1, 3, 5).reshape(t * oh * ow, d * ph * pw)


def _merge_tokens(y: torch.Tensor, t: int, d: int, H: int, W: int, pw: int, ph: int) -> torch.Tensor:
    """inverse of _split_tokens (auto_sttn.py:201-202)."""
    oh, ow = H // ph, W // pw
    return y.reshape(t, oh, ow, d, ph, pw).permute(0, 3, 1, 4, 2, 5).reshape(t, d, H, W)


def multihead_patch_attention(w, prefix: str, x: torch.Tensor, patchsize=PATCHSIZE) -> torch.Tensor:
    """MultiHeadedAttention.forward, auto_sttn.py:167-206, with b = 1 (infer :111-115)."""
    t, c, H, W = x.shape
    q = F.conv2d(x, w[prefix + "query_embedding.weight"], w[prefix + "query_embedding.bias"])
    k = F.conv2d(x, w[prefix + "key_embedding.weight"], w[prefix + "key_embedding.bias"])
    v = F.conv2d(x, w[prefix + "value_embedding.weight"], w[prefix + "value_embedding.bias"])
    dk = c // len(patchsize)
    outs = []
    for i, (pw, ph) in enumerate(patchsize):
        sl = slice(i * dk, (i + 1) * dk)
        qt, kt, vt = (_split_tokens(z[:, sl], pw, ph) for z in (q, k, v))
        scores = (qt @ kt.t()) / math.sqrt(qt.shape[-1])  # Attention.forward :141-142
        p = torch.softmax(scores, dim=-1)
        outs.append(_merge_tokens(p @ vt, t, dk, H, W, pw, ph))
    y = torch.cat(outs, dim=1)
    return _lrelu(F.conv2d(y, w[prefix + "output_linear.0.weight"], w[prefix + "output_linear.0.bias"], padding=1))


def feed_forward(w, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """FeedForward, auto_sttn.py:210-222: 3x3 dilation 2 pad 2 -> LReLU -> 3x3 pad 1 -> LReLU."""
    x = _lrelu(F.conv2d(x, w[prefix + "conv.0.weight"], w[prefix + "conv.0.bias"], padding=2, dilation=2))
    return _lrelu(F.conv2d(x, w[prefix + "conv.2.weight"], w[prefix + "conv.2.bias"], padding=1))


def transformer_block(w, b: int, x: torch.Tensor, patchsize=PATCHSIZE) -> torch.Tensor:
    """TransformerBlock.forward, auto_sttn.py:235-239."""
    p = f"transformer.{b}."
    x = x + multihead_patch_attention(w, p + "attention.", x, patchsize)
    return x + feed_forward(w, p + "feed_forward.", x)


def infer(w, feat: torch.Tensor, patchsize=PATCHSIZE, taps: Optional[list] = None) -> torch.Tensor:
    """InpaintGenerator.infer, auto_sttn.py:111-115 (8 blocks).  `taps` collects per-block outputs."""
    x = feat
    for b in range(STACK_NUM):
        x = transformer_block(w, b, x, patchsize)
        if taps is not None:
            taps.append(x)
    return x


def _up2(x):
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)  # deconv :124-126


def decoder(w, x: torch.Tensor) -> torch.Tensor:
    """auto_sttn.py:87-95 (no tanh here; the caller applies it, sttn_auto_inpaint.py:150)."""
    x = _lrelu(F.conv2d(_up2(x), w["decoder.0.conv.weight"], w["decoder.0.conv.bias"], padding=1))
    x = _lrelu(F.conv2d(x, w["decoder.2.weight"], w["decoder.2.bias"], padding=1))
    x = _lrelu(F.conv2d(_up2(x), w["decoder.4.conv.weight"], w["decoder.4.conv.bias"], padding=1))
    return F.conv2d(x, w["decoder.6.weight"], w["decoder.6.bias"], padding=1)


# ----------------------------------------------------------------------------------------------
# A4 / A11: strip-level inpaint
# ----------------------------------------------------------------------------------------------
def frames_to_tensor(frames_bgr: Sequence[np.ndarray]) -> torch.Tensor:
    """Stack + ToTorchFormatTensor + `*2-1` (utils/sttn_utils.py:66-112, sttn_auto_inpaint.py:128):
    BGR u8 HWC -> RGB float [T,3,H,W] in [-1,1]."""
    arr = np.stack([f[:, :, ::-1] for f in frames_bgr]).astype(np.float32)  # T,H,W,3 RGB
    x = torch.from_numpy(arr).permute(0, 3, 1, 2).contiguous()
    return x.div(255) * 2 - 1


def quantise(pred_img: torch.Tensor) -> np.ndarray:
    """sttn_auto_inpaint.py:150-158: tanh -> (x+1)/2 -> *255 (fp32) -> astype(uint8) (truncation)."""
    y = (torch.tanh(pred_img) + 1) / 2
    return (y.permute(0, 2, 3, 1).numpy() * 255).astype(np.uint8)


def inpaint_strip(w, frames_bgr: Sequence[np.ndarray], stride: int = NEIGHBOR_STRIDE, ref_length: int = REF_LENGTH,
                  patchsize=PATCHSIZE, taps: Optional[dict] = None) -> List[np.ndarray]:
    """STTNInpaint.inpaint, sttn_auto_inpaint.py:122-164.  frames: T x [120,640,3] u8 BGR.
    Returns T comps: RGB, uint8 if the frame was decoded once, else float32 (0.5/0.5 running blend)."""
    T = len(frames_bgr)
    with torch.no_grad():
        feats = encoder(w, frames_to_tensor(frames_bgr))
        if taps is not None:
            taps["encoder"] = feats
        comps: List[Optional[np.ndarray]] = [None] * T
        for wi, (nb, refs) in enumerate(window_schedule(T, stride, ref_length)):
            pred = infer(w, feats[nb + refs], patchsize)
            img = quantise(decoder(w, pred[:len(nb)]))
            if taps is not None and wi == 0:
                taps["window0_feat"] = pred
                taps["window0_img"] = img
            for i, idx in enumerate(nb):
                if comps[idx] is None:
                    comps[idx] = img[i]
                else:
                    comps[idx] = comps[idx].astype(np.float32) * 0.5 + img[i].astype(np.float32) * 0.5
    return comps  # type: ignore[return-value]


def upscale_comp(comp: np.ndarray, W: int, split_h: int) -> np.ndarray:
    """sttn_auto_inpaint.py:86-87 / :312-313: cv2.resize to (W, split_h), astype(uint8), swap R<->B."""
    if comp.dtype == np.uint8:
        up = cv2_resize_linear_u8(comp, W, split_h)
    else:
        up = cv2_resize_linear_f32(comp, W, split_h).astype(np.uint8)
    return up[:, :, ::-1]


# ----------------------------------------------------------------------------------------------
# A12: the drop-in call
# ----------------------------------------------------------------------------------------------
def sttn_call(w, input_frames: Sequence[np.ndarray], input_mask: np.ndarray, stride: int = NEIGHBOR_STRIDE,
              ref_length: int = REF_LENGTH) -> List[np.ndarray]:
    """STTNInpaint.__call__, sttn_auto_inpaint.py:43-97.  Frames BGR u8 [H,W,3]; mask u8 [H,W] 0/255."""
    mask = (input_mask > 127).astype(np.uint8)[:, :, None]  # cv2.threshold(…,127,1,THRESH_BINARY) :48
    H, W = mask.shape[:2]
    split_h = int(W * 3 / 16)  # :54
    areas = get_inpaint_area_by_mask(W, H, split_h, mask)
    frames = [f.copy() for f in input_frames]
    if not areas:
        return frames
    comps = {}
    for k, (y0, y1, _, _) in enumerate(areas):
        scaled = [cv2_resize_linear_u8(np.ascontiguousarray(f[y0:y1]), MODEL_W, MODEL_H) for f in frames]
        comps[k] = inpaint_strip(w, scaled, stride, ref_length)
    for j, frame in enumerate(frames):
        for k, (y0, y1, _, _) in enumerate(areas):
            comp = upscale_comp(comps[k][j], W, split_h)
            m = mask[y0:y1]
            frame[y0:y1] = m * comp + (1 - m) * frame[y0:y1]  # :91
    return frames


def sttn_video(w, frames: Sequence[np.ndarray], input_mask: np.ndarray, clip_gap: int = MAX_LOAD_NUM) -> List[np.ndarray]:
    """STTNAutoInpaint.__call__ chunk loop, sttn_auto_inpaint.py:242-328, on an in-memory clip with
    ab_sections=None: independent chunks of exactly clip_gap frames (last one shorter)."""
    out: List[np.ndarray] = []
    for s in range(0, len(frames), clip_gap):
        out.extend(sttn_call(w, frames[s:s + clip_gap], input_mask))
    return out


def psnr_u8(a: np.ndarray, b: np.ndarray) -> float:
    """backend/inpaint/video/core/metrics.py:36 — 20*log10(255/sqrt(mse))."""
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else 20 * math.log10(255.0 / math.sqrt(mse))


# ----------------------------------------------------------------------------------------------
# synthetic clip of SURVEY.md §8(d)
# ----------------------------------------------------------------------------------------------
def synthetic_clip(n: int, H: int, W: int, seed: int = 0, pad: int = 64) -> List[np.ndarray]:
    """Blurred-noise background translating (3,2) px/frame; deterministic in (n,H,W,seed).
    Pure numpy (separable box blur x3 ≈ Gaussian) so the GPU box needs no cv2 for the generator."""
    rng = np.random.default_rng(seed)
    bh, bw = H + 2 * pad, W + 3 * pad
    bg = rng.integers(0, 256, (bh // 4 + 2, bw // 4 + 2, 3), dtype=np.uint8).astype(np.float32)
    # cheap smooth upsample x4 by bilinear, then one box blur
    yy = np.linspace(0, bg.shape[0] - 1.001, bh).astype(np.float32)
    xx = np.linspace(0, bg.shape[1] - 1.001, bw).astype(np.float32)
    y0 = np.floor(yy).astype(np.int64); x0 = np.floor(xx).astype(np.int64)
    fy = (yy - y0)[:, None, None]; fx = (xx - x0)[None, :, None]
    big = (bg[y0][:, x0] * (1 - fy) * (1 - fx) + bg[y0][:, x0 + 1] * (1 - fy) * fx
           + bg[y0 + 1][:, x0] * fy * (1 - fx) + bg[y0 + 1][:, x0 + 1] * fy * fx)
    big = np.clip(big, 0, 255).astype(np.uint8)
    frames = []
    for i in range(n):
        oy = (2 * i) % (2 * pad)
        ox = (3 * i) % (3 * pad)
        frames.append(np.ascontiguousarray(big[oy:oy + H, ox:ox + W]))
    return frames


def default_mask(H: int, W: int, area=(0.88, 0.99, 0.15, 0.85)) -> np.ndarray:
    """backend/config.py:43 default selection area "ymin,ymax,xmin,xmax" fractions -> create_mask."""
    ymin, ymax, xmin, xmax = int(H * area[0]), int(H * area[1]), int(W * area[2]), int(W * area[3])
    return create_mask((H, W), [(xmin, xmax, ymin, ymax)])
