"""GPU: the DBNet text detector compiled from the reference's PIR program against the oracle
interpreter of the same program (oracle/dbnet_oracle.py).  Forward parity is pinned to the reference's model files;
box post-processing is 'parity unpinned' (no paddleocr here) and compared against the restatement.
Tolerance: the device network multiplies in fp16 (fp32 accumulate) through ~150 layers: |prob diff| <= 0.06 max,
<= 2e-3 mean; boxes within 3 px."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import dbnet_oracle as D

pytestmark = pytest.mark.gpu
MODEL_DIR = os.path.join(ROOT, "weights", "V5", "ch_det")


def _text_frame(h=720, w=1280, seed=0):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(seed)
    img = cv2.GaussianBlur(rng.integers(0, 255, (h, w, 3), dtype=np.uint8), (0, 0), 9)
    for i, (txt, y) in enumerate((("The quick brown fox 0123", int(h * 0.9)), ("SUBTITLE line two", int(h * 0.5)))):
        x = int(w * 0.2) + 40 * i
        cv2.putText(img, txt, (x, y), cv2.FONT_HERSHEY_SIMPLEX, 1.6, (0, 0, 0), 8, cv2.LINE_AA)
        cv2.putText(img, txt, (x, y), cv2.FONT_HERSHEY_SIMPLEX, 1.6, (255, 255, 255), 3, cv2.LINE_AA)
    return img


@pytest.fixture(scope="module")
def detector(capi):
    if not os.path.exists(os.path.join(MODEL_DIR, "inference.pdiparams")):
        pytest.skip("detector model not staged under weights/V5/ch_det")
    if capi.lib().vsr_device_count() < 1:
        pytest.fail("GPU tests need a B200 (sm_100) device")
    from vsr_b200.dbnet import TextDetector

    return TextDetector(MODEL_DIR, "cuda:0"), D.Graph(MODEL_DIR)


def test_probability_map_matches_oracle(detector):
    det, graph = detector
    img = _text_frame()
    got = det.probability_map(img)
    x = D.preprocess(img)
    want = D.forward(graph, x)[0, 0].numpy()
    assert got.shape == want.shape == (544, 960)
    d = np.abs(got - want)
    assert np.isfinite(got).all()
    assert d.mean() <= 2e-3 and d.max() <= 0.06, (float(d.mean()), float(d.max()))
    assert ((got > 0.3) != (want > 0.3)).mean() < 2e-3  # the binarised map the boxes come from


def test_boxes_and_filtering(detector):
    det, graph = detector
    img = _text_frame(seed=1)
    polys = det.predict(img)[0]["dt_polys"]
    want = D.postprocess(D.forward(graph, D.preprocess(img))[0, 0].numpy(), img.shape[0], img.shape[1])
    assert len(polys) == len(want) >= 2
    a = sorted(D.get_coordinates(polys.tolist()))
    b = sorted(D.get_coordinates(want.tolist()))
    assert all(max(abs(p - q) for p, q in zip(x, y)) <= 3 for x, y in zip(a, b)), (a, b)
    from vsr_b200 import SubtitleDetect

    sd = SubtitleDetect("", sub_areas=[(int(720 * 0.8), 719, 0, 1279)], model_dir=MODEL_DIR)
    sd._detector = det
    boxes = sd.detect_subtitle(img)
    assert len(boxes) == 1 and boxes[0][2] >= int(720 * 0.8)  # only the bottom line is inside the selected area
