"""CPU: the host-side pieces of the detector (model-file parsing, resize rule, DB post-process) against the oracle;
the oracle's forward pass on the reference's real program produces a sane text map (pins the interpreter to the
model files — SURVEY §8c: there is no paddleocr here to compare with)."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import dbnet_oracle as D

MODEL_DIR = os.path.join(ROOT, "weights", "V5", "ch_det")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(MODEL_DIR, "inference.pdiparams")),
                                reason="detector model not staged under weights/V5/ch_det")


def test_model_files_parse_identically():
    from vsr_b200 import dbnet

    nodes, params = dbnet._load_program(MODEL_DIR)
    g = D.Graph(MODEL_DIR)
    assert len(nodes) == len(g.ops) == 1052 and len(params) == len(g.params) == 539
    for k, v in g.params.items():
        assert np.array_equal(params[k], v)
    kinds = {}
    for n in nodes:
        kinds[n.kind] = kinds.get(n.kind, 0) + 1
    assert kinds["conv2d"] == 115 and kinds["depthwise_conv2d"] == 27 and kinds["batch_norm_"] == 87


def test_resize_rule():
    from vsr_b200.dbnet import TextDetector

    for hw, want in (((1080, 1920), (544, 960)), ((720, 1280), (544, 960)), ((2160, 3840), (544, 960)), ((480, 852), (480, 864)),
                     ((1280, 720), (960, 544)), ((20, 20), (32, 32))):
        assert TextDetector.resize_shape(*hw) == want == D.resize_shape(*hw)


@pytest.mark.slow
def test_oracle_forward_and_postprocess():
    cv2 = pytest.importorskip("cv2")
    from vsr_b200.dbnet import db_postprocess

    img = np.full((360, 640, 3), 90, np.uint8)
    cv2.putText(img, "HELLO WORLD 42", (120, 320), cv2.FONT_HERSHEY_SIMPLEX, 1.2, (255, 255, 255), 3, cv2.LINE_AA)
    prob = D.forward(D.Graph(MODEL_DIR), D.preprocess(img))[0, 0].numpy()
    assert prob.shape == (352, 640) and 0.0 <= prob.min() and prob.max() <= 1.0
    want = D.postprocess(prob, 360, 640)
    got = db_postprocess(prob, 360, 640)
    assert np.array_equal(got, want) and len(want) == 1
    (xmin, xmax, ymin, ymax), = D.get_coordinates(want.tolist())
    assert 90 <= xmin <= 125 and 400 <= xmax <= 460 and 270 <= ymin <= 300 and 320 <= ymax <= 345


@pytest.mark.parametrize("hw", [(96, 160), (64, 224)])
def test_graph_compiler_on_cpu_runtime(hw):
    """The PIR -> runtime compiler (BN / bias / ReLU folding, concat layout and channel permutations, SAME padding,
    stride-2 and transposed convs) driven on a CPU stand-in of the device runtime equals the oracle interpreter."""
    import cv2
    from fake_rt import FakeRuntime
    from vsr_b200.dbnet import TextDetector

    rng = np.random.default_rng(5)
    img = rng.integers(0, 255, hw + (3,), dtype=np.uint8)
    cv2.putText(img, "Ab3", (10, hw[0] - 20), cv2.FONT_HERSHEY_SIMPLEX, 1.5, (255, 255, 255), 3)
    rt = FakeRuntime()
    det = TextDetector(MODEL_DIR, runtime=rt)
    got = det.probability_map(img)           # first call: calibration of the tensor scales, layer by layer
    want = D.forward(D.Graph(MODEL_DIR), D.preprocess(img))[0, 0].numpy()
    assert got.shape == want.shape == hw
    assert np.abs(got - want).max() < 2e-4
    prog = det._programs[hw]
    scales = sorted({t.scale for t in prog.values.values()})
    assert rt.rescaled > 0 and scales[0] < 1.0 and scales[-1] == 1.0   # the neck does not fit fp16 unscaled
    assert all(float(np.log2(s)).is_integer() for s in scales)
    n0 = rt.launch_count
    again = det.probability_map(img)         # second call: the recorded graph
    assert np.array_equal(again, got)
    assert rt.launch_count - n0 < 330        # 1052 PIR ops -> one launch per fused conv / add / pool / concat part
    # scales that are too large for a frame (here: tampered with) overflow fp16 in the recorded graph -> the flag of the
    # scaled epilogues -> recalibration on that frame, still the right answer
    for t in prog.values.values():
        if t.follow is None and t.scale < 1.0:
            t.scale = min(t.scale * 4096.0, 1.0)
    rt.capture_begin()
    for st in prog.steps:
        st.run(rt)
    prog.graph = rt.capture_end()
    n1 = rt.launch_count
    third = det.probability_map(img)
    assert rt.launch_count - n1 > 400        # the graph run plus the layer-by-layer recalibration
    assert np.abs(third - want).max() < 2e-4


def test_mobile_detector_compiles_on_cpu_runtime():
    """PP-OCRv5_mobile_det (backend/models/V5/ch_det_fast, model_config.py:17-18): re-parameterised PPLCNetV3 blocks (conv + bias
    + learnable scalar affine folded into one launch, hardswish + affine as one element-wise launch), squeeze-and-excitation
    as a fused gate, RSEFPN with residual gates, 42/18/12-channel convs padded to the 8-channel store granularity."""
    import cv2
    from fake_rt import FakeRuntime
    from vsr_b200.dbnet import TextDetector

    d = os.path.join(ROOT, "weights", "V5", "ch_det_fast")
    if not os.path.exists(os.path.join(d, "inference.pdiparams")):
        pytest.skip("mobile detector not staged under weights/V5/ch_det_fast")
    rng = np.random.default_rng(6)
    img = rng.integers(0, 255, (96, 160, 3), dtype=np.uint8)
    cv2.putText(img, "Ab3", (10, 76), cv2.FONT_HERSHEY_SIMPLEX, 1.5, (255, 255, 255), 3)
    rt = FakeRuntime()
    det = TextDetector(d, runtime=rt)
    got = det.probability_map(img)
    want = D.forward(D.Graph(d), D.preprocess(img))[0, 0].numpy()
    assert got.shape == want.shape == (96, 160) and np.abs(got - want).max() < 2e-4
    n0 = rt.launch_count
    assert np.array_equal(det.probability_map(img), got) and rt.launch_count - n0 < 130     # 683 PIR ops


def test_batched_detection_equals_single_frames_on_cpu_runtime():
    """`probability_maps`: several sampled frames in one launch give each frame's own result (calibration sees the whole batch)."""
    import cv2
    from fake_rt import FakeRuntime
    from vsr_b200.dbnet import TextDetector

    rng = np.random.default_rng(9)
    imgs = []
    for i in range(3):
        img = rng.integers(0, 255, (64, 160, 3), dtype=np.uint8)
        cv2.putText(img, f"t{i}", (10 + 30 * i, 50), cv2.FONT_HERSHEY_SIMPLEX, 1.2, (255, 255, 255), 3)
        imgs.append(img)
    det = TextDetector(MODEL_DIR, runtime=FakeRuntime())
    maps = det.probability_maps(imgs)
    g = D.Graph(MODEL_DIR)
    for m, img in zip(maps, imgs):
        assert np.abs(m - D.forward(g, D.preprocess(img))[0, 0].numpy()).max() < 2e-4
    assert np.abs(det.probability_map(imgs[1]) - maps[1]).max() < 1e-5
    assert sorted(k for k in det._programs if len(k) == 3) == [(3, 64, 160)]


def test_weight_scaling_and_split_weights_in_fp16_simulation():
    """The detector's fp16 parity used to be limited by fp16 UNDERFLOW of the tiny weights behind the un-normalised neck (max |dp| ~0.24 on
    this frame in the stand-in's fp16 simulation, ~0.3 measured on the B200).  `_weight_scale` packs such layers multiplied by a power of
    two and divides it out in the epilogue: max |dp| < 0.1, mean 4x smaller.  Keeping the weights as hi + lo fp16 halves
    (`precise_weights`, opt-in) takes out the remaining weight rounding."""
    import cv2
    from fake_rt import FakeRuntime
    from vsr_b200 import dbnet
    from vsr_b200.dbnet import TextDetector

    rng = np.random.default_rng(0)
    img = cv2.GaussianBlur(rng.integers(0, 255, (192, 320, 3), dtype=np.uint8), (0, 0), 9)
    cv2.putText(img, "quick brown", (30, 150), cv2.FONT_HERSHEY_SIMPLEX, 1.2, (255, 255, 255), 3, cv2.LINE_AA)
    want = D.forward(D.Graph(MODEL_DIR), D.preprocess(img))[0, 0].numpy()

    def err(**kw):
        with np.errstate(over="ignore"):
            return np.abs(TextDetector(MODEL_DIR, runtime=FakeRuntime(fp16=True), **kw).probability_map(img) - want)

    scaled, split = err(precise_weights=False), err(precise_weights=True)
    saved = dbnet._weight_scale
    dbnet._weight_scale = lambda w: 1.0
    try:
        unscaled = err(precise_weights=False)
    finally:
        dbnet._weight_scale = saved
    assert unscaled.max() > 0.15 and scaled.max() < 0.1 and scaled.mean() < unscaled.mean() / 2
    assert split.max() <= scaled.max() and split.mean() <= scaled.mean() * 1.05
    assert dbnet._weight_scale(np.array([3e-7, -1e-8], np.float32)) == 2.0 ** 22 and dbnet._weight_scale(np.array([0.3], np.float32)) == 1.0


def test_batched_mobile_detector_gates_each_image_on_cpu_runtime():
    """The mobile model's squeeze-and-excitation gates are per-image statistics: in a batch they are computed image by image (a dark and
    a bright frame in one launch keep their own gates)."""
    import cv2
    from fake_rt import FakeRuntime
    from vsr_b200.dbnet import TextDetector

    d = os.path.join(ROOT, "weights", "V5", "ch_det_fast")
    if not os.path.exists(os.path.join(d, "inference.pdiparams")):
        pytest.skip("mobile detector not staged under weights/V5/ch_det_fast")
    rng = np.random.default_rng(9)
    imgs = []
    for i in range(2):
        img = rng.integers(0, 255, (64, 160, 3), dtype=np.uint8) // (1 + 3 * i)
        cv2.putText(img, f"t{i}", (10 + 30 * i, 50), cv2.FONT_HERSHEY_SIMPLEX, 1.2, (255, 255, 255), 3)
        imgs.append(img)
    det = TextDetector(d, runtime=FakeRuntime())
    g = D.Graph(d)
    for m, img in zip(det.probability_maps(imgs), imgs):
        assert np.abs(m - D.forward(g, D.preprocess(img))[0, 0].numpy()).max() < 2e-4
