#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the reference tree is not on the GPU box):
    python tools/make_golden.py
Fixtures (all inputs are rebuilt from seeds by oracle.sttn_oracle.synthetic_clip, so only outputs and
small taps are stored):
  sttn_auto_strip_real.npz   STTNInpaint.inpaint() on 7 strip frames [120,640,3], real weights
  sttn_auto_call_real.npz    STTNInpaint.__call__() on 8 frames 800x450 + default mask, real weights
  sttn_auto_strip_rand.npz   same network class loaded with oracle.random_weights(seed=0)
  sttn_det_real.npz          STTNDetInpaint.inpaint() + __call__() on 7 frames 640x360, real sttn-det weights
  mask_index.npz             create_mask / get_inpaint_area_by_mask / batch_generator vectors
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, sttn_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def comps_pack(comps):
    """comps are u8 (visited once) or f32 (blended): store as f32 + a visited-once flag."""
    return (np.stack([c.astype(np.float32) for c in comps]), np.array([c.dtype == np.uint8 for c in comps]))


def main():
    ref_import.install()
    import cv2
    from backend.inpaint.sttn_auto_inpaint import STTNInpaint
    from backend.tools import inpaint_tools as RT

    os.makedirs(OUT, exist_ok=True)
    dev = torch.device("cpu")
    real = STTNInpaint(dev, ref_import.weights_path("sttn-auto"))

    # ---- strip level, real weights
    T = 7
    strip = [cv2.resize(f, (640, 120)) for f in O.synthetic_clip(T, 360, 1920, seed=11)]
    taps = {}

    def hook(name):
        def h(mod, inp, out):
            o = out["x"] if isinstance(out, dict) else out
            taps.setdefault(name, o.detach().clone())
        return h

    hs = [real.model.encoder.register_forward_hook(hook("encoder")),
          real.model.transformer[0].register_forward_hook(hook("block0")),
          real.model.transformer[7].register_forward_hook(hook("block7"))]
    comps = real.inpaint([s.copy() for s in strip])
    for h in hs:
        h.remove()
    c, once = comps_pack(comps)
    np.savez_compressed(os.path.join(OUT, "sttn_auto_strip_real.npz"),
                        seed=11, T=T, comps=c, once=once,
                        encoder_f0=taps["encoder"][0, :, ::3, ::8].numpy().astype(np.float32),
                        block0_f0=taps["block0"][0, :, ::3, ::8].numpy().astype(np.float32),
                        block7_f0=taps["block7"][0, :, ::3, ::8].numpy().astype(np.float32))

    # ---- full call, real weights (IPP-enabled cv2, as a user runs it) + the plain-C++ cv2 path
    H, W, T = 450, 800, 8
    frames = O.synthetic_clip(T, H, W, seed=3)
    mask = O.default_mask(H, W)
    out = real([f.copy() for f in frames], mask)
    cv2.setUseOptimized(False)
    out_plain = real([f.copy() for f in frames], mask)
    cv2.setUseOptimized(True)
    areas = RT.get_inpaint_area_by_mask(W, H, int(W * 3 / 16), (mask > 127).astype(np.uint8)[:, :, None])
    y0, y1 = areas[0][:2]
    untouched = all(np.array_equal(o[:y0], f[:y0]) and np.array_equal(o[y1:], f[y1:]) for o, f in zip(out, frames))
    np.savez_compressed(os.path.join(OUT, "sttn_auto_call_real.npz"), seed=3, H=H, W=W, T=T, areas=np.array(areas),
                        untouched_outside=untouched,
                        strip_out=np.stack([o[y0:y1] for o in out]),
                        plain_diff_idx=np.flatnonzero(np.stack(out) != np.stack(out_plain)).astype(np.int64),
                        plain_diff_val=np.stack(out_plain).ravel()[np.flatnonzero(np.stack(out) != np.stack(out_plain))])

    # ---- strip level, seeded random weights in the reference network class
    from backend.inpaint.sttn.auto_sttn import InpaintGenerator
    rnd = STTNInpaint.__new__(STTNInpaint)
    rnd.device = dev
    rnd.model = InpaintGenerator(init_weights=False)
    rnd.model.load_state_dict(O.random_weights(0))
    rnd.model.eval()
    rnd.model_input_width, rnd.model_input_height = 640, 120
    rnd.neighbor_stride, rnd.ref_length = 5, 10
    T = 6
    strip = [cv2.resize(f, (640, 120)) for f in O.synthetic_clip(T, 360, 1920, seed=12)]
    c, once = comps_pack(rnd.inpaint([s.copy() for s in strip]))
    np.savez_compressed(os.path.join(OUT, "sttn_auto_strip_rand.npz"), seed=12, T=T, wseed=0, comps=c, once=once)

    # ---- sttn-det (D1-D3): strip level + full call, real weights
    from backend.inpaint.sttn_det_inpaint import STTNDetInpaint
    from oracle import sttn_det_oracle as D

    det = STTNDetInpaint(dev, ref_import.weights_path("sttn-det"))
    H, W, T = 360, 640, 7
    frames = O.synthetic_clip(T, H, W, seed=9)
    mask = O.default_mask(H, W)
    sh = D.split_height(H, W)
    areas = RT.get_inpaint_area_by_mask(W, H, sh, mask[:, :, None])
    y0, y1 = areas[0][:2]
    scaled = [cv2.resize(f[y0:y1], (432, 240)) for f in frames]
    msmall = cv2.resize(mask[y0:y1], (432, 240))
    c, once = comps_pack(det.inpaint([s_.copy() for s_ in scaled], [msmall.copy() for _ in scaled]))
    cv2.setUseOptimized(False)
    out_plain = det([f.copy() for f in frames], mask)
    cv2.setUseOptimized(True)
    np.savez_compressed(os.path.join(OUT, "sttn_det_real.npz"), seed=9, H=H, W=W, T=T, areas=np.array(areas), comps=c, once=once,
                        mask_small=msmall, strip_out_plain=np.stack([o[y0:y1] for o in out_plain]))

    # ---- T1/T3/T4 planning functions of SubtitleDetect + expand_frame_ranges
    from backend.tools.ocr import get_coordinates as ref_get_coordinates
    from backend.tools.subtitle_detect import SubtitleDetect
    import json as _json

    prng = np.random.default_rng(8)
    sd = SubtitleDetect.__new__(SubtitleDetect)
    plan = []
    for _ in range(60):
        step = int(prng.choice([2, 3, 4]))
        n = int(prng.integers(20, 200))
        sampled = {}
        for f in range(1, n + 1, step):
            if prng.random() < 0.6:
                base = (int(prng.integers(100, 140)), int(prng.integers(500, 560)), int(prng.integers(400, 420)), int(prng.integers(440, 470)))
                boxes = [base]
                if prng.random() < 0.3:
                    boxes.append((base[0] + 5, base[1] - 5, base[2] - 60, base[3] - 60))
                sampled[f] = boxes
        if not sampled:
            continue
        # phase 2 of find_subtitle_frame_no (subtitle_detect.py:112-132), verbatim calls into the reference helpers
        d, nos = {}, sorted(sampled)
        for a, b in zip(nos, nos[1:]):
            d[a] = sampled[a]
            if b - a <= step * 2:
                for ff in range(a + 1, b):
                    d[ff] = sampled[a]
        d[nos[-1]] = sampled[nos[-1]]
        uni = sd.unify_regions(dict(d))
        uni = {k: v for k, v in uni.items() if len(v) > 0}
        ranges = SubtitleDetect.find_continuous_ranges_with_same_mask(uni)
        ranges0 = SubtitleDetect.find_continuous_ranges(uni)
        expanded = RT.expand_frame_ranges(ranges, 3, 3)
        merged = SubtitleDetect.filter_and_merge_intervals(expanded, 10)
        pts = sorted(int(x) for x in prng.integers(1, n, size=int(prng.integers(0, 5))))
        split = SubtitleDetect.split_range_by_scene(list(merged), list(pts))
        plan.append(dict(step=step, sampled={str(k): v for k, v in sampled.items()}, filled={str(k): v for k, v in d.items()},
                         unified={str(k): v for k, v in uni.items()}, ranges=ranges, ranges0=ranges0, expanded=expanded, merged=merged,
                         points=pts, split=split))
    quads = prng.integers(0, 1000, size=(30, 4, 2)).tolist()
    np.savez_compressed(os.path.join(OUT, "subtitle_plan.npz"), plan=_json.dumps(plan), quads=_json.dumps(quads),
                        coords=_json.dumps(ref_get_coordinates(quads)))

    # ---- integer path vectors
    rng = np.random.default_rng(5)
    cases = []
    for _ in range(40):
        Hh, Ww = int(rng.integers(60, 400)), int(rng.integers(80, 640))
        boxes = []
        for _ in range(int(rng.integers(1, 4))):
            x0 = int(rng.integers(0, Ww - 10)); x1 = int(rng.integers(x0 + 1, Ww))
            y0_ = int(rng.integers(0, Hh - 5)); y1_ = int(rng.integers(y0_ + 1, Hh))
            boxes.append((x0, x1, y0_, y1_))
        m = RT.create_mask((Hh, Ww), boxes)
        h = int(Ww * 3 / 16)
        a1 = RT.get_inpaint_area_by_mask(Ww, Hh, h, (m > 127).astype(np.uint8)[:, :, None])
        a8 = RT.get_inpaint_area_by_mask(Ww, Hh, h, (m > 127).astype(np.uint8)[:, :, None], multiple=8)
        cases.append(dict(H=Hh, W=Ww, boxes=boxes, mask_sum=int(m.astype(np.int64).sum()),
                          mask_rows=np.flatnonzero(m.any(1))[[0, -1]].tolist(), areas=a1, areas8=a8))
    bg = {f"{n}@{mb}": [len(b) for b in RT.batch_generator(list(range(n)), mb)]
          for n in (300, 50, 500, 1200, 299, 1, 7, 49, 100) for mb in (50, 70, 3)}
    import json
    np.savez_compressed(os.path.join(OUT, "mask_index.npz"), cases=json.dumps(cases), batches=json.dumps(bg))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
