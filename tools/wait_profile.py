#!/usr/bin/env python
"""Where do the warp roles of the tcgen05 kernels wait?  Needs the profiling build:
    nvcc ... -DVSR_TC_PROFILE engine.cu -o csrc/build/libvsr_b200_prof.so
    VSR_B200_LIB=.../libvsr_b200_prof.so python tools/wait_profile.py
Prints, per kernel, the average SM cycles per CTA spent waiting in each role."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import sttn_oracle as O  # noqa: E402
from vsr_b200 import STTNInpaint, _capi, ops  # noqa: E402

NAMES = {0: "conv256 1-CTA", 1: "conv256 2-CTA pair", 2: "scores", 3: "pv", 4: "conv BN<256", 5: "scores 2-CTA", 6: "pv 2-CTA"}
SLOTS = ["producer waits empty", "mma waits operands", "mma waits accumulator", "epilogue waits accumulator", "epilogue busy",
         "kernel cycles"]


def dump(tag):
    buf = (C.c_uint64 * 64)()
    _capi.check(_capi.lib().vsr_debug_tc_profile(buf, 1))
    print(f"--- {tag}")
    for kid, name in NAMES.items():
        n = buf[kid * 8 + 6]
        if not n:
            continue
        vals = [buf[kid * 8 + s] / n for s in range(6)]
        tot = vals[5]
        print(f"  {name:20s} CTAs={n:6d} " + "  ".join(f"{SLOTS[i]}={vals[i]:9.0f} ({100 * vals[i] / tot:4.1f}%)" for i in range(6)))


eng = STTNInpaint("cuda:0", {k: v.numpy() for k, v in O.random_weights(0).items()})
dump("warm-up (ignored)")
eng.time_conv(15, 10)
dump(f"conv 3x3 256->256, T=15 (VSR_CONV_2CTA={os.environ.get('VSR_CONV_2CTA', '1')})")
rng = np.random.default_rng(0)
q, k, v = (rng.standard_normal((15, 30, 160, 256), dtype=np.float32) for _ in range(3))
ops.patch_attention(q, k, v, [(80, 15), (32, 6), (10, 5), (5, 3)])
dump("patch attention T=15, 4 heads")
for hd in ([(5, 3)], [(10, 5)], [(32, 6)], [(80, 15)]):
    ops.patch_attention(q[..., :64], k[..., :64], v[..., :64], hd)
    dump(f"patch attention T=15, head {hd}")
