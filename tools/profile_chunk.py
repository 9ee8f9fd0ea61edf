#!/usr/bin/env python
"""Two 1080p 50-frame chunks through the engine (first = warm-up) — the command wrapped by ncu for the
launch list and the --set full captures committed under profiles/."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import sttn_oracle as O  # noqa: E402
from vsr_b200 import STTNInpaint  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 50
p = os.path.join(ROOT, "weights", "sttn-auto", "infer_model.pth")
w = p if os.path.exists(p) else {k: v.numpy() for k, v in O.random_weights(0).items()}
eng = STTNInpaint("cuda:0", w)
frames = O.synthetic_clip(T, 1080, 1920, seed=0)
mask = O.default_mask(1080, 1920)
for i in range(2):
    n0 = eng.launch_count
    eng.inpaint_inplace([f.copy() for f in frames], mask)
    print(f"chunk {i}: {eng.launch_count - n0} kernel launches", flush=True)
