"""Layer-by-layer comparison of the device detector with the oracle interpreter (run on a GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
from oracle import dbnet_oracle as D
from vsr_b200.dbnet import TextDetector

MODEL = os.path.join("weights", "V5", "ch_det")
rng = np.random.default_rng(0)
h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (720, 1280)
img = cv2.GaussianBlur(rng.integers(0, 255, (h, w, 3), dtype=np.uint8), (0, 0), 9)
cv2.putText(img, "The quick brown fox 0123", (int(w * .2), int(h * .9)), cv2.FONT_HERSHEY_SIMPLEX, 1.6, (255, 255, 255), 3, cv2.LINE_AA)
det = TextDetector(MODEL, "cuda:0")
got = det.probability_map(img)
vals = {}
want = D.forward(D.Graph(MODEL), D.preprocess(img), values=vals)[0, 0].numpy()
prog = next(iter(det._programs.values()))
print("final: max diff", np.abs(got - want).max(), "nonfinite", (~np.isfinite(got)).sum())
bad = 0
for vid in sorted(prog.values):
    t = prog.values[vid]
    if vid not in vals:
        continue
    ref = vals[vid][0].permute(1, 2, 0).numpy()
    dev = det._rt.download(t).astype(np.float32) / t.scale
    idx = t.perm if t.perm is not None else np.arange(t.c)
    dev = dev[:, :, idx]
    nf = int((~np.isfinite(dev)).sum())
    d = np.abs(np.nan_to_num(dev, posinf=1e9, neginf=-1e9) - ref)
    scale = np.abs(ref).max()
    flag = "  <<<<" if (nf or d.max() > 0.05 * max(scale, 1.0)) else ""
    if flag:
        bad += 1
    if flag or "-v" in sys.argv:
        print(f"id {vid:5d} c={t.c:4d} {t.h}x{t.w} cp={t.cp} ref|max|={scale:9.3f} maxdiff={d.max():9.4f} mean={d.mean():.5f} nonfinite={nf}{flag}")
    if bad >= 12:
        break
