"""CPU: the LAMA oracle (oracle/lama_oracle.py) pinned against the unmodified reference: the golden
outputs of `LamaInpaint.inpaint` / `__call__` (tools/make_golden_lama.py) and, when the TorchScript file is staged, the
script module itself."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
from oracle import lama_oracle as L
from oracle import sttn_oracle as O

sys.path.insert(0, os.path.join(ROOT, "tools"))
PT = os.path.join(ROOT, "weights", "big-lama", "big-lama.pt")
needs_weights = pytest.mark.skipif(not os.path.exists(PT), reason="big-lama.pt not staged under weights/big-lama")


def test_shapes_and_padding():
    s = L.weight_shapes()
    assert sum(int(np.prod(v)) for v in s.values()) == 51_057_027   # the script module: 51,057,179 including its 152 batch counters
    a = np.arange(2 * 5 * 7, dtype=np.float32).reshape(2, 5, 7)
    p = L.pad_img_to_modulo(a, 8)    # symmetric: the edge sample is repeated first
    assert p.shape == (2, 8, 8) and np.array_equal(p[:, 5, :7], a[:, 4]) and np.array_equal(p[:, 7, :7], a[:, 2])
    assert np.array_equal(p[:, :5, 7], a[:, :, 6])
    g = L.get_image(np.full((3, 4, 3), 51, np.uint8))
    assert g.shape == (3, 3, 4) and g.dtype == np.float32 and np.all(g == np.float32(51) / np.float32(255))


@needs_weights
def test_forward_equals_torchscript():
    w = L.load_weights(PT)
    m = torch.jit.load(PT, map_location="cpu").eval()
    rng = np.random.default_rng(0)
    img = torch.from_numpy(rng.random((2, 3, 64, 104), dtype=np.float32))
    mask = torch.zeros(2, 1, 64, 104)
    mask[:, :, 20:40, 30:80] = 1
    with torch.inference_mode():
        ref = m(img, mask)
    assert torch.equal(L.forward(w, img, mask), ref)


@needs_weights
def test_golden_reference_outputs():
    from make_golden_lama import inputs

    w = L.load_weights(PT)
    z = np.load(os.path.join(GOLDEN, "lama_real.npz"))
    img, m, frames, mask = inputs()
    assert np.array_equal(L.inpaint(w, img, m), z["single"])
    out = L.lama_call(w, frames, mask)
    assert np.array_equal(np.stack(out), z["call"])
    assert not np.array_equal(np.stack(out), np.stack(frames))


@needs_weights
def test_config1_golden_reference_outputs():
    """BASELINE config 1 (tests/golden/config1_lama.npz, tools/make_golden_configs.py): the oracle equals the unmodified reference's
    `LamaInpaint.inpaint` on the 512x512 synthetic image and on frame 0 of test/test.mp4 + test/test.png, byte for byte."""
    w = L.load_weights(PT)
    z = np.load(os.path.join(GOLDEN, "config1_lama.npz"))
    assert np.array_equal(L.inpaint(w, z["test_frame0"], z["test_mask"]), z["test_out"])
    img = O.synthetic_clip(1, 512, 512, seed=0)[0]
    m = np.zeros((512, 512), np.uint8)
    m[400:470, 60:450] = 255
    assert np.array_equal(L.inpaint(w, img, m), z["out512"])
