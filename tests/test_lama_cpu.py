"""CPU: the LAMA graph builder (vsr_b200.lama_inpaint.LamaNetwork: folded batch-norms, fused local/global tensor,
channel-slice views, reflect padding through padded-grid convs with cropped stores, stride-2 and transposed convs as dense
convs, FourierUnit layout) driven on the fp32 stand-in of the device runtime (tests/fake_rt.py) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import lama_oracle as L


@pytest.mark.parametrize("hw", [(45, 70), (64, 48)])
def test_lama_builder_on_cpu_runtime(hw):
    from fake_rt import FakeRuntime
    from vsr_b200.lama_inpaint import LamaInpaint

    w = L.random_weights(3)
    rt = FakeRuntime()
    eng = LamaInpaint("cuda:0", {k: v.numpy() for k, v in w.items()}, runtime=rt)
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, hw + (3,), dtype=np.uint8)
    mask = np.zeros(hw, np.uint8)
    mask[hw[0] // 3: 2 * hw[0] // 3, 10: hw[1] - 8] = 255
    got = eng.inpaint(img, mask)
    want = L.inpaint(w, img, mask)
    assert got.shape == want.shape == hw + (3,) and got.dtype == np.uint8
    d = np.abs(got.astype(np.int32) - want)
    assert np.array_equal(got[mask == 0], img[mask == 0]) or d[mask == 0].max() <= 1
    assert d.max() <= 1 and (d > 0).mean() < 0.01            # fp32 re-association only
    assert np.abs(got[mask > 0].astype(np.int32) - img[mask > 0]).mean() > 5   # the hole was really repainted
    n0 = rt.launch_count
    again = eng.inpaint(img, mask)                            # second call replays the recorded graph
    assert np.array_equal(again, got) and rt.launch_count - n0 < 560   # 18 blocks x (2 x 14 + 1) launches + stem, downs, ups, head


def test_lama_batches_and_strip_call():
    """`__call__` on 6 frames: strips go through the network 4 + 2 per launch (two compiled batch sizes); every frame equals
    its own single-image result, and the oracle's `lama_call`."""
    from fake_rt import FakeRuntime
    from oracle import sttn_oracle as O
    from vsr_b200.lama_inpaint import LamaInpaint

    w = L.random_weights(4)
    eng = LamaInpaint("cuda:0", {k: v.numpy() for k, v in w.items()}, runtime=FakeRuntime())
    H, W = 128, 160
    frames = O.synthetic_clip(6, H, W, seed=41)
    keep = [f.copy() for f in frames]
    mask = O.create_mask((H, W), [(30, 130, 96, 118)])
    out = eng(frames, mask)
    assert all(np.array_equal(a, b) for a, b in zip(frames, keep))
    assert sorted(k[0] for k in eng.model._programs) == [2, 4]
    want = L.lama_call(w, frames, mask)
    for o, r in zip(out, want):
        d = np.abs(o.astype(np.int32) - r)
        assert d.max() <= 1 and (d > 0).mean() < 0.01
    (y0, y1, _, _), = O.get_inpaint_area_by_mask(W, H, int(W * 3 / 16), mask)
    single = eng.inpaint(frames[4][y0:y1], mask[y0:y1])
    assert np.abs(single.astype(np.int32) - out[4][y0:y1]).max() <= 1


def test_fp16_storage_simulation_forecasts_gpu_parity():
    """The stand-in's fp16 mode (fp16 tensor storage + fp16 tensor-core operands, fp32 accumulation / FFT / residual master) predicts
    the GPU parity: on the 70x100 golden case it gives ~55.6 dB / max 2 where the B200 measured 57.4 dB / max 1 (profiles/lama_r1.json).
    Used to forecast paths that have not run on a GPU yet (profiles/fp16_forecast_r1.md, tests/diag_fp16_forecast.py)."""
    import os
    import sys

    from conftest import ROOT
    from fake_rt import FakeRuntime
    from oracle import sttn_oracle as O
    from vsr_b200.lama_inpaint import LamaInpaint

    pt = os.path.join(ROOT, "weights", "big-lama", "big-lama.pt")
    if not os.path.exists(pt):
        pytest.skip("big-lama.pt not staged")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_golden_lama import inputs

    w = L.load_weights(pt)
    img, m = inputs()[:2]
    got = LamaInpaint("cuda:0", {k: v.numpy() for k, v in w.items()}, runtime=FakeRuntime(fp16=True)).inpaint(img, m)
    want = L.inpaint(w, img, m)
    hole = m > 0
    assert np.array_equal(got[~hole], want[~hole])
    assert O.psnr_u8(got[hole].astype(np.float32), want[hole].astype(np.float32)) >= 50.0 and np.abs(got.astype(np.int32) - want).max() <= 4
