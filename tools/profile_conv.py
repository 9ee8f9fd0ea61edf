#!/usr/bin/env python
"""The exact launch bench.py's `roofline` times (3x3 conv 256->256 on 29x30x160 px, LeakyReLU, fp16 out),
a few times in a row — the command wrapped by `ncu --set full` for profiles/ncu_*_conv.*"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sttn_oracle as O  # noqa: E402
from vsr_b200 import STTNInpaint  # noqa: E402

eng = STTNInpaint("cuda:0", {k: v.numpy() for k, v in O.random_weights(0).items()})
T = int(sys.argv[1]) if len(sys.argv) > 1 else 29
ms = eng.time_conv(T, 6)
print("conv ms per launch:", [round(float(m), 4) for m in ms])
