#!/usr/bin/env python
"""Goldens of the UNMODIFIED reference at the sizes of the BASELINE.json configurations (build container only; the reference tree
and the TorchScript big-lama do not travel).  Inputs are the seeded clips of bench.py, rebuilt by the tests; only outputs are stored,
cut down to what a small fixture can hold (the runs themselves are whole):

  config2_sttn_auto_1080p.npz   STTNInpaint.__call__ (sttn_auto_inpaint.py:43-97) on ONE WHOLE 50-frame 1920x1080 chunk (synthetic_clip seed 0,
                                default-bbox mask = the chunk bench.py times): frames FRAMES, the bounding box of the mask; per-frame sums of
                                all 50 output strips
  config4_sttn_det_1080p.npz    STTNDetInpaint.__call__ (sttn_det_inpaint.py:38-99) on one 46-frame 1080p batch (seed 100): frames FRAMES of the
                                533-row strip, every second row and column; per-frame sums of all 46 strips
  config1_lama.npz              LamaInpaint.inpaint (lama_inpaint.py:17-28) on the 512x512 synthetic image of SURVEY §8d (texture seed 0, hole rows
                                400-470, cols 60-450) and on frame 0 of test/test.mp4 with test/test.png (input frame and mask stored: 852x480)

    python tools/make_golden_configs.py [2] [4] [1]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, sttn_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
FRAMES2 = [0, 5, 24, 25, 45, 49]
FRAMES4 = [0, 23, 45]


def config2():
    from backend.inpaint.sttn_auto_inpaint import STTNInpaint

    H, W, T = 1080, 1920, 50
    frames = O.synthetic_clip(T, H, W, seed=0)
    mask = O.default_mask(H, W)
    ys, xs = np.nonzero(mask)
    box = (int(ys.min()), int(ys.max()) + 1, int(xs.min()), int(xs.max()) + 1)
    t0 = time.time()
    out = STTNInpaint(torch.device("cpu"), ref_import.weights_path("sttn-auto"))([f.copy() for f in frames], mask)
    print(f"config 2: {T} frames in {time.time() - t0:.0f} s")
    keep = np.stack(out)[:, :720]
    assert np.array_equal(keep, np.stack(frames)[:, :720])           # rows above the strip untouched
    np.savez_compressed(os.path.join(OUT, "config2_sttn_auto_1080p.npz"), seed=0, H=H, W=W, T=T, frames=np.array(FRAMES2), box=np.array(box),
                        out_box=np.stack([out[i][box[0]:box[1], box[2]:box[3]] for i in FRAMES2]),
                        strip_sums=np.array([int(o[720:].astype(np.int64).sum()) for o in out]))


def config4():
    from backend.inpaint.sttn_det_inpaint import STTNDetInpaint

    H, W, T = 1080, 1920, 46
    frames = O.synthetic_clip(T, H, W, seed=100)
    mask = O.default_mask(H, W)
    t0 = time.time()
    out = STTNDetInpaint(torch.device("cpu"), ref_import.weights_path("sttn-det"))([f.copy() for f in frames], mask)
    print(f"config 4: {T} frames in {time.time() - t0:.0f} s")
    diff = np.flatnonzero((np.stack(out) != np.stack(frames)).any(axis=(0, 2, 3)))
    y0, y1 = int(diff.min()), int(diff.max()) + 1
    np.savez_compressed(os.path.join(OUT, "config4_sttn_det_1080p.npz"), seed=100, H=H, W=W, T=T, frames=np.array(FRAMES4), rows=np.array([y0, y1]),
                        out_half=np.stack([out[i][y0:y1:2, ::2] for i in FRAMES4]),
                        strip_sums=np.array([int(o[y0:y1].astype(np.int64).sum()) for o in out]))


def config1():
    import cv2
    from backend.inpaint.lama_inpaint import LamaInpaint

    model = LamaInpaint(torch.device("cpu"), os.path.join(ROOT, "weights", "big-lama", "big-lama.pt"))
    img = O.synthetic_clip(1, 512, 512, seed=0)[0]
    m = np.zeros((512, 512), np.uint8)
    m[400:470, 60:450] = 255
    t0 = time.time()
    out512 = model.inpaint(img.copy(), m.copy())
    cap = cv2.VideoCapture(os.path.join(ref_import.REF_ROOT, "test", "test.mp4"))
    ok, frame0 = cap.read()
    assert ok
    tm = cv2.imread(os.path.join(ref_import.REF_ROOT, "test", "test.png"), 0)
    out_t = model.inpaint(frame0.copy(), tm.copy())
    print(f"config 1: two images in {time.time() - t0:.0f} s", out512.shape, out_t.shape)
    np.savez_compressed(os.path.join(OUT, "config1_lama.npz"), out512=out512, test_frame0=frame0, test_mask=tm, test_out=out_t)


def main():
    ref_import.install()
    torch.manual_seed(0)
    which = sys.argv[1:] or ["2", "4", "1"]
    if "1" in which:
        config1()
    if "4" in which:
        config4()
    if "2" in which:
        config2()
    for f in ("config1_lama.npz", "config2_sttn_auto_1080p.npz", "config4_sttn_det_1080p.npz"):
        p = os.path.join(OUT, f)
        if os.path.exists(p):
            print(f, os.path.getsize(p))


if __name__ == "__main__":
    main()
