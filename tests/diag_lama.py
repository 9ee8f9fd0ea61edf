"""LAMA on the device against the oracle with the reference's weights, plus timings (diagnostic script for a GPU box:
`python tests/diag_lama.py`; not collected by pytest)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import lama_oracle as L, sttn_oracle as O
from vsr_b200.lama_inpaint import LamaInpaint

path = os.path.join(ROOT, "weights", "big-lama", "big-lama.npz")
w = L.load_weights(path)
eng = LamaInpaint("cuda:0", path)
res = {}
for name, (H, W) in (("small", (120, 256)), ("strip1080p", (360, 1920))):
    img = O.synthetic_clip(1, H, W, seed=5)[0]
    mask = np.zeros((H, W), np.uint8)
    mask[int(H * .55):int(H * .85), int(W * .15):int(W * .85)] = 255
    got = eng.inpaint(img, mask)
    t0 = time.perf_counter(); want = L.inpaint(w, img, mask); cpu_s = time.perf_counter() - t0
    hole = mask > 0
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    res[name] = {"psnr_hole": float(O.psnr_u8(got[hole].astype(np.float32), want[hole].astype(np.float32))), "max": int(d.max()),
                 "mean_abs_hole": float(d[hole].mean()), "outside_exact": bool(np.array_equal(got[~hole], want[~hole])), "oracle_cpu_s": cpu_s}
    N = 10
    t0 = time.perf_counter()
    for _ in range(N):
        eng.inpaint(img, mask)
    res[name]["e2e_ms"] = (time.perf_counter() - t0) / N * 1e3
    res[name]["network_ms"] = eng.model.time_network(10)
    t0 = time.perf_counter()
    for _ in range(3):
        eng._inpaint_batch([img] * 8, [mask] * 8)
    res[name]["batch4_e2e_ms_per_frame"] = (time.perf_counter() - t0) / 24 * 1e3
    res[name]["batch4_network_ms_per_frame"] = eng.model.time_network(10) / 4
print(json.dumps(res, indent=1))
