#!/usr/bin/env python
"""Generate tests/golden/lama_real.npz by running the UNMODIFIED reference LamaInpaint (/root/reference) on CPU with the
reference's big-lama TorchScript (its 5 parts concatenated in fs_manifest.csv order -> weights/big-lama/big-lama.pt,
tools/stage_weights.py).  Build container only.  Inputs are rebuilt from seeds by the tests; stored: outputs.

  single   LamaInpaint.inpaint(image 70x100x3, mask 70x100)          (pads to 72x104, crops back)
  call     LamaInpaint.__call__(3 frames 120x256x3, create_mask box) (strip height int(256*3/16) = 48, batch path)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, sttn_oracle as O  # noqa: E402


def inputs():
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (70, 100, 3), dtype=np.uint8)
    m = np.zeros((70, 100), np.uint8)
    m[30:52, 20:83] = 255
    frames = O.synthetic_clip(3, 120, 256, seed=12)
    mask = O.create_mask((120, 256), [(40, 215, 80, 108)])
    return img, m, frames, mask


def main():
    ref_import.install()
    from backend.inpaint.lama_inpaint import LamaInpaint

    torch.manual_seed(0)
    model = LamaInpaint(torch.device("cpu"), os.path.join(ROOT, "weights", "big-lama", "big-lama.pt"))
    img, m, frames, mask = inputs()
    single = model.inpaint(img, m)
    out = model([f.copy() for f in frames], mask)
    p = os.path.join(ROOT, "tests", "golden", "lama_real.npz")
    np.savez_compressed(p, single=single, call=np.stack(out))
    print(p, os.path.getsize(p), single.shape, np.stack(out).shape)


if __name__ == "__main__":
    main()
