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
