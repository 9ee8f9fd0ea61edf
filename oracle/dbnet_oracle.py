"""ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE) — CPU restatement of the text-detection pass
(SURVEY.md §8a rows T2/T3): PP-OCRv5 DBNet forward + DB post-process + the reference's box filtering.

The arithmetic lives in third-party `paddleocr==3.4.0` / `paddlepaddle==3.0.0` (requirements.txt:9), neither of
which is vendored under /root/reference nor installable here.  What the reference DOES pin is the network:
`backend/models/V5/ch_det/inference.json` (Paddle PIR program) + `inference.pdiparams` + `inference.yml`
(pre/post-processing constants).  This module executes that program with torch ops (op semantics of SURVEY
Appendix A.6) and restates PaddleX's DetResizeForTest / NormalizeImage / DBPostProcess from their published
algorithm (Appendix A.5).

PARITY STATUS: the forward pass is pinned to the reference's model files (same graph, same weights); the
pre/post-processing is **parity unpinned** — no paddleocr install, golden vector or test of the reference
exists to check it against.  Box-level parity must therefore be read as "against this restatement".
Reference call sites: backend/tools/subtitle_detect.py:43-82, backend/tools/ocr.py:1-20.
"""
from __future__ import annotations

import json
import math
import os
import struct
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# backend/models/V5/ch_det/inference.yml
RESIZE_LONG = 960
MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)
THRESH, BOX_THRESH, MAX_CANDIDATES, UNCLIP_RATIO = 0.3, 0.6, 1000, 1.5


# ------------------------------------------------------------------------------------------------ model files
def read_pdiparams(path: str) -> List[np.ndarray]:
    """Stream of {u32 version, u64 lod_level, [lod], u32 tensor version, i32 desc_len, TensorDesc proto, raw}
    records; tensors come in sorted(parameter name) order."""
    b = open(path, "rb").read()
    off, out = 0, []

    def varint(buf, j):
        v = sh = 0
        while True:
            c = buf[j]
            j += 1
            v |= (c & 0x7F) << sh
            sh += 7
            if c < 0x80:
                return v, j

    while off < len(b):
        off += 4
        (lod,) = struct.unpack_from("<Q", b, off)
        off += 8
        for _ in range(lod):
            (sz,) = struct.unpack_from("<Q", b, off)
            off += 8 + sz
        off += 4
        (dlen,) = struct.unpack_from("<i", b, off)
        off += 4
        desc = b[off:off + dlen]
        off += dlen
        j, dims, dtype = 0, [], None
        while j < len(desc):
            tag = desc[j]
            j += 1
            if tag & 7 == 0:
                v, j = varint(desc, j)
                if tag >> 3 == 1:
                    dtype = v
                else:
                    dims.append(v)
            elif tag & 7 == 2:
                ln, j = varint(desc, j)
                end = j + ln
                while j < end:
                    v, j = varint(desc, j)
                    dims.append(v)
        assert dtype == 5, f"only fp32 tensors expected, got proto dtype {dtype}"
        n = int(np.prod(dims)) if dims else 1
        out.append(np.frombuffer(b, dtype="<f4", count=n, offset=off).reshape(dims).copy())
        off += 4 * n
    return out


def _attr(a):
    t = a["AT"]
    d = t.get("D")
    if t["#"] == "0.a_array":
        return [x["D"] for x in d]
    return d


class Graph:
    """The PIR program: ops in execution order + parameters by name."""

    def __init__(self, model_dir: str):
        g = json.load(open(os.path.join(model_dir, "inference.json")))
        self.ops = g["program"]["regions"][0]["blocks"][0]["ops"]
        names = sorted(o["A"][3] for o in self.ops if o["#"] == "p")
        tensors = read_pdiparams(os.path.join(model_dir, "inference.pdiparams"))
        assert len(names) == len(tensors), (len(names), len(tensors))
        self.params: Dict[str, np.ndarray] = dict(zip(names, tensors))
        for o in self.ops:
            if o["#"] == "p":
                shape = o["O"]["TT"]["D"][1]
                assert list(self.params[o["A"][3]].shape) == list(shape), o["A"][3]


def _same_pad(n, k, s, d=1):
    out = -(-n // s)
    return max((out - 1) * s + (k - 1) * d + 1 - n, 0)


def forward(graph: Graph, x: torch.Tensor, taps: Dict[str, torch.Tensor] | None = None, values: Dict[int, torch.Tensor] | None = None) -> torch.Tensor:
    """Execute the program on x [N,3,H,W] fp32 -> probability map [N,1,H,W] (op table: SURVEY A.6).
    `values`, when given, receives every activation by PIR value id (layer-by-layer diagnosis of the device path)."""
    v: Dict[int, object] = {}
    with torch.no_grad():
        for o in graph.ops:
            kind = o["#"]
            ins = [v.get(i["%"]) for i in o.get("I", [])]
            at = {a["N"]: _attr(a) for a in o.get("A", []) if isinstance(a, dict)}
            outs = o.get("O", [])
            oid = outs["%"] if isinstance(outs, dict) else (outs[0]["%"] if outs else None)
            if kind == "p":
                r = torch.from_numpy(graph.params[o["A"][3]])
            elif kind == "1.data":
                r = x
            elif kind in ("1.conv2d", "1.depthwise_conv2d"):
                t, w = ins
                pad = at["paddings"]
                if at["padding_algorithm"] == "SAME":
                    ph = _same_pad(t.shape[2], w.shape[2], at["strides"][0], at["dilations"][0])
                    pw = _same_pad(t.shape[3], w.shape[3], at["strides"][1], at["dilations"][1])
                    t = F.pad(t, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
                    pad = [0, 0]
                r = F.conv2d(t, w, None, at["strides"], pad, at["dilations"], at["groups"])
            elif kind == "1.conv2d_transpose":
                r = F.conv_transpose2d(ins[0], ins[1], None, at["strides"], at["paddings"], 0, at["groups"], at["dilations"])
            elif kind == "1.batch_norm_":
                t, mean, var, scale, bias = ins
                r = F.batch_norm(t, mean, var, scale, bias, False, 0.0, at["epsilon"])
            elif kind == "1.relu":
                r = F.relu(ins[0])
            elif kind == "1.sigmoid":
                r = torch.sigmoid(ins[0])
            elif kind == "1.hardswish":
                r = F.hardswish(ins[0])
            elif kind == "1.hardsigmoid":
                r = torch.clamp(ins[0] * at["slope"] + at["offset"], 0, 1)
            elif kind == "1.add":
                r = ins[0] + ins[1]
            elif kind == "1.multiply":
                r = ins[0] * ins[1]
            elif kind == "1.full_int_array":
                r = list(at["value"])
            elif kind == "1.full":
                r = at["value"]
            elif kind == "0.combine":
                r = list(ins)
            elif kind == "1.concat":
                r = torch.cat(ins[0], int(ins[1]))
            elif kind == "1.reshape":
                r = ins[0].reshape([int(d) for d in ins[1]])
            elif kind == "1.nearest_interp":
                sc = at.get("scale") or [2.0, 2.0]  # 2, 4 or 8 in the FPN heads (align_corners False)
                r = F.interpolate(ins[0], scale_factor=(float(sc[0]), float(sc[1])), mode="nearest")
            elif kind == "1.pool2d":
                t, ks = ins[0], [int(k) for k in ins[1]]
                if at.get("adaptive"):
                    r = F.adaptive_avg_pool2d(t, ks) if at["pooling_type"] == "avg" else F.adaptive_max_pool2d(t, ks)
                else:
                    st, pad = at["strides"], at["paddings"]
                    if at["padding_algorithm"] == "SAME":
                        ph = _same_pad(t.shape[2], ks[0], st[0])
                        pw = _same_pad(t.shape[3], ks[1], st[1])
                        t = F.pad(t, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2), value=float("-inf") if at["pooling_type"] == "max" else 0.0)
                        pad = [0, 0]
                    fn = F.max_pool2d if at["pooling_type"] == "max" else F.avg_pool2d
                    r = fn(t, ks, st, pad[:2], ceil_mode=bool(at.get("ceil_mode", False)))
            elif kind == "1.scale":
                r = ins[0] * float(ins[1]) + float(at.get("bias", 0.0))
            elif kind == "1.fetch":
                return ins[0]
            else:
                raise NotImplementedError(kind)
            if oid is not None:
                v[oid] = r
                if values is not None and isinstance(r, torch.Tensor) and r.dim() == 4 and kind != "p":
                    values[oid] = r
                if taps is not None and isinstance(r, torch.Tensor) and "struct_name" in at and kind.startswith("1.conv"):
                    taps[f"{oid}:{at['struct_name']}"] = r
    raise RuntimeError("program has no fetch op")


# ------------------------------------------------------------------------------------------------ pre / post
def resize_shape(h: int, w: int, limit: int = RESIZE_LONG) -> Tuple[int, int]:
    """DetResizeForTest(resize_long=960) -> limit_side_len 960, limit_type 'max' (unpinned):
    shrink so the long side is <= 960, then round each side to a multiple of 32 (min 32)."""
    ratio = limit / max(h, w) if max(h, w) > limit else 1.0
    rh, rw = int(h * ratio), int(w * ratio)
    rh = max(int(round(rh / 32) * 32), 32)
    rw = max(int(round(rw / 32) * 32), 32)
    return rh, rw


def preprocess(img_bgr: np.ndarray) -> torch.Tensor:
    """BGR u8 HWC -> [1,3,rh,rw] fp32: cv2.resize INTER_LINEAR, x/255, (x-mean)/std applied in BGR order with the
    ImageNet constants as listed (inference.yml:22-40), HWC -> CHW."""
    from . import sttn_oracle as O

    rh, rw = resize_shape(*img_bgr.shape[:2])
    r = O.cv2_resize_linear_u8(np.ascontiguousarray(img_bgr), rw, rh).astype(np.float32)
    r = (r * np.float32(1.0 / 255.0) - np.asarray(MEAN, np.float32)) / np.asarray(STD, np.float32)
    return torch.from_numpy(np.ascontiguousarray(r.transpose(2, 0, 1)))[None]


def _order_quad(pts: np.ndarray) -> np.ndarray:
    """get_mini_boxes: sort by x, split the pairs by y -> [top-left, top-right, bottom-right, bottom-left]."""
    p = sorted(pts.tolist(), key=lambda q: q[0])
    (i1, i4) = (0, 1) if p[1][1] > p[0][1] else (1, 0)
    (i2, i3) = (2, 3) if p[3][1] > p[2][1] else (3, 2)
    return np.array([p[i1], p[i2], p[i3], p[i4]], dtype=np.float32)


def postprocess(prob: np.ndarray, src_h: int, src_w: int) -> np.ndarray:
    """DBPostProcess(thresh 0.3, box_thresh 0.6, max_candidates 1000, unclip 1.5, quad boxes, fast score) as
    published in PaddleOCR — **unpinned**.  The Clipper offset of the min-area rectangle is
    restated as growing that rectangle by d = area*ratio/perimeter on every side (equal up to Clipper's
    integer rounding).  Returns int16 quads [N,4,2] in source-image pixels."""
    import cv2

    H, W = prob.shape
    bitmap = (prob > THRESH).astype(np.uint8)
    contours, _ = cv2.findContours(bitmap * 255, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    boxes = []
    for c in contours[:MAX_CANDIDATES]:
        rect = cv2.minAreaRect(c)
        if min(rect[1]) < 3:
            continue
        quad = _order_quad(cv2.boxPoints(rect))
        # box_score_fast: mean probability inside the quad
        xs, ys = quad[:, 0], quad[:, 1]
        x0, x1 = int(np.clip(np.floor(xs.min()), 0, W - 1)), int(np.clip(np.ceil(xs.max()), 0, W - 1))
        y0, y1 = int(np.clip(np.floor(ys.min()), 0, H - 1)), int(np.clip(np.ceil(ys.max()), 0, H - 1))
        m = np.zeros((y1 - y0 + 1, x1 - x0 + 1), np.uint8)
        cv2.fillPoly(m, [(quad - np.array([x0, y0], np.float32)).astype(np.int32)], 1)
        score = cv2.mean(prob[y0:y1 + 1, x0:x1 + 1], m)[0]
        if score < BOX_THRESH:
            continue
        (cx, cy), (rw, rh), ang = rect
        d = (rw * rh) * UNCLIP_RATIO / (2 * (rw + rh))
        grown = ((cx, cy), (rw + 2 * d, rh + 2 * d), ang)
        if min(grown[1]) < 5:
            continue
        q = _order_quad(cv2.boxPoints(grown))
        q[:, 0] = np.clip(np.round(q[:, 0] / W * src_w), 0, src_w)
        q[:, 1] = np.clip(np.round(q[:, 1] / H * src_h), 0, src_h)
        boxes.append(q.astype(np.int16))
    return np.array(boxes, dtype=np.int16).reshape(-1, 4, 2)


# ------------------------------------------------------------------------------------------------ reference glue
def get_coordinates(dt_polys: Sequence) -> List[Tuple[int, int, int, int]]:
    """backend/tools/ocr.py:1-20: quad -> (xmin, xmax, ymin, ymax) = (max(x1,x4), min(x2,x3), max(y1,y2), min(y3,y4))."""
    out = []
    for q in dt_polys:
        (x1, y1), (x2, y2), (x3, y3), (x4, y4) = [(int(p[0]), int(p[1])) for p in q]
        out.append((max(x1, x4), min(x2, x3), max(y1, y2), min(y3, y4)))
    return out


def filter_boxes(coords, sub_areas) -> List[Tuple[int, int, int, int]]:
    """backend/tools/subtitle_detect.py:60-82: keep boxes lying fully inside one of the selected areas
    (ymin, ymax, xmin, xmax); no areas -> keep all."""
    if not sub_areas:
        return list(coords)
    keep = []
    for xmin, xmax, ymin, ymax in coords:
        for s_ymin, s_ymax, s_xmin, s_xmax in sub_areas:
            if s_xmin <= xmin and xmax <= s_xmax and s_ymin <= ymin and ymax <= s_ymax:
                keep.append((xmin, xmax, ymin, ymax))
                break
    return keep


def detect_subtitle(graph: Graph, img_bgr: np.ndarray, sub_areas=None):
    """SubtitleDetect.detect_subtitle (subtitle_detect.py:56-82) with TextDetection.predict restated."""
    prob = forward(graph, preprocess(img_bgr))[0, 0].numpy()
    polys = postprocess(prob, img_bgr.shape[0], img_bgr.shape[1])
    return filter_boxes(get_coordinates(polys.tolist()), sub_areas)


def sample_step(fps: float) -> int:
    """subtitle_detect.py:29-39."""
    return 4 if fps >= 60 else 3 if fps >= 30 else 2
