"""Time the B200 text detector on one 1080p frame (device network + host post-process), print a JSON line."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsr_b200.dbnet import TextDetector
from vsr_b200 import _capi

det = TextDetector(os.path.join("weights", "V5", "ch_det"), "cuda:0")
import cv2
yy, xx = np.mgrid[0:1080, 0:1920].astype(np.float32)
img = np.stack([96 + 60 * np.sin(xx / 211 + c) + 50 * np.cos(yy / 173 - c) for c in range(3)], -1).clip(0, 255).astype(np.uint8)
for txt, org in (("The quick brown fox 0123", (400, 1000)), ("second line of a subtitle", (500, 930))):
    cv2.putText(img, txt, org, cv2.FONT_HERSHEY_SIMPLEX, 2.0, (0, 0, 0), 9, cv2.LINE_AA)
    cv2.putText(img, txt, org, cv2.FONT_HERSHEY_SIMPLEX, 2.0, (255, 255, 255), 4, cv2.LINE_AA)
for _ in range(3):
    det.probability_map(img)
n0 = det.launch_count
t0 = time.perf_counter()
N = 20
for _ in range(N):
    p = det.probability_map(img)
t1 = time.perf_counter()
launches = (det.launch_count - n0) / N
from vsr_b200.dbnet import db_postprocess
t2 = time.perf_counter()
for _ in range(N):
    r = det.predict(img)
t3 = time.perf_counter()
t4 = time.perf_counter()
for _ in range(N):
    db_postprocess(p, 1080, 1920)
t5 = time.perf_counter()
print(json.dumps({"probability_map_ms": (t1 - t0) / N * 1e3, "predict_ms": (t3 - t2) / N * 1e3, "launches_per_frame": launches,
                  "boxes": len(r[0]["dt_polys"]), "postprocess_ms": (t5 - t4) / N * 1e3, "frames_per_s": N / (t3 - t2)}))
