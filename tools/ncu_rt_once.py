"""One eager pass of the detector or of LAMA for an ncu launch list (`ncu --metrics gpu__time_duration.sum ... python tools/ncu_rt_once.py lama`)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
which = sys.argv[1] if len(sys.argv) > 1 else "lama"
H, W = 1080, 1920
yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
img = np.stack([96 + 60 * np.sin(xx / 211 + c) + 50 * np.cos(yy / 173 - c) for c in range(3)], -1).clip(0, 255).astype(np.uint8)
if which == "dbnet":
    from vsr_b200.dbnet import TextDetector
    TextDetector(os.path.join("weights", "V5", "ch_det"), "cuda:0").probability_map(img)
else:
    from vsr_b200 import LamaInpaint
    strip = np.ascontiguousarray(img[700:1060])
    mask = np.zeros((360, 1920), np.uint8)
    mask[200:320, 300:1600] = 255
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    LamaInpaint("cuda:0", os.path.join("weights", "big-lama", "big-lama.npz"))._inpaint_batch([strip] * n, [mask] * n)
