"""fp16-storage simulation of the device paths on the CPU stand-in: forecast of the GPU parity (compare with measured where known)."""
import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/tools')
import numpy as np, torch
from fake_rt import FakeRuntime
from oracle import sttn_oracle as O
which = sys.argv[1]
if which == "lama":
    from oracle import lama_oracle as L
    from vsr_b200.lama_inpaint import LamaInpaint
    from make_golden_lama import inputs
    w = L.load_weights('/root/repo/weights/big-lama/big-lama.pt')
    eng = LamaInpaint("cuda:0", {k: v.numpy() for k, v in w.items()}, runtime=FakeRuntime(fp16=True))
    img, m, frames, mask = inputs()
    got = eng.inpaint(img, m); want = L.inpaint(w, img, m)
    hole = m > 0
    d = np.abs(got.astype(int) - want)
    print("LAMA 70x100 sim: psnr", O.psnr_u8(got[hole].astype(np.float32), want[hole].astype(np.float32)), "max", d.max(), " (GPU measured: 57.4 dB / max 1)")
elif which == "dbnet":
    import cv2
    from oracle import dbnet_oracle as D
    from vsr_b200.dbnet import TextDetector
    d = '/root/repo/weights/V5/ch_det'
    rng = np.random.default_rng(0)
    img = cv2.GaussianBlur(rng.integers(0, 255, (360, 640, 3), dtype=np.uint8), (0, 0), 9)
    cv2.putText(img, "The quick brown fox", (100, 320), cv2.FONT_HERSHEY_SIMPLEX, 1.2, (255, 255, 255), 3, cv2.LINE_AA)
    det = TextDetector(d, runtime=FakeRuntime(fp16=True))
    got = det.probability_map(img); want = D.forward(D.Graph(d), D.preprocess(img))[0, 0].numpy()
    dd = np.abs(got - want)
    print("DBNet 352x640 sim: mean", dd.mean(), "frac>0.05", (dd > 0.05).mean(), "max", dd.max(), " (GPU measured: mean 1.8e-4, frac 6e-4)")
elif which == "propainter":
    from make_golden_propainter import inputs
    from vsr_b200.propainter_inpaint import PropainterInpaint
    from oracle import propainter_oracle as P
    z = np.load('/root/repo/tests/golden/propainter_real.npz')
    frames, mask = inputs()[:2]
    t = time.time()
    out = np.stack(PropainterInpaint("cuda:0", '/root/repo/weights/propainter', runtime=FakeRuntime(fp16=True)).inpaint(frames, mask))
    _, md = P.read_mask(mask, len(frames))
    hole = np.stack(md) > 0
    d = np.abs(out.astype(int) - z["comp"])
    print("ProPainter sim (%.0f s): psnr hole" % (time.time() - t), O.psnr_u8(out[hole].astype(np.float32), z["comp"][hole].astype(np.float32)), "max", d.max(), "frac>2", (d > 2).mean(),
          "outside exact", bool(np.array_equal(out[~hole], np.stack(frames)[~hole])))
