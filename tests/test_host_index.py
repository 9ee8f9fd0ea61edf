"""CPU: the product's host C++ integer path (through the C ABI) against the oracle, property-style, plus
that the library loads and exports every symbol include/vsr_b200.h declares."""
import os
import re

import numpy as np
import pytest

from oracle import sttn_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(capi):
    L = capi.lib()
    hdr = open(os.path.join(ROOT, "include", "vsr_b200.h")).read()
    declared = set(re.findall(r"\b(vsr_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/vsr_b200.h but not exported"
    assert declared == set(capi.EXPORTED_SYMBOLS)
    assert b"sm_100a" in L.vsr_version()


def test_compute_fails_loudly_without_gpu(capi):
    import ctypes as C

    L = capi.lib()
    if L.vsr_device_count() > 0:
        pytest.skip("a B200 is present")
    h = C.c_void_p()
    rc = L.vsr_sttn_create(C.byref(h), 0, None)
    assert rc < 0 and b"no CPU fallback" in L.vsr_last_error()
    from vsr_b200 import STTNInpaint

    with pytest.raises(capi.VsrError):
        STTNInpaint("cuda:0", {k: v.numpy() for k, v in O.random_weights(0).items()})
    with pytest.raises(capi.VsrError):
        STTNInpaint("cpu", {})


def test_create_mask_matches_oracle(capi):
    from vsr_b200 import create_mask

    rng = np.random.default_rng(3)
    for _ in range(200):
        H, W = int(rng.integers(8, 300)), int(rng.integers(8, 400))
        boxes = []
        for _ in range(int(rng.integers(0, 5))):
            x0 = int(rng.integers(-5, W)); x1 = int(rng.integers(x0, W + 30))
            y0 = int(rng.integers(-5, H)); y1 = int(rng.integers(y0, H + 30))
            boxes.append((x0, x1, y0, y1))
        assert np.array_equal(create_mask((H, W), boxes), O.create_mask((H, W), boxes))


def test_inpaint_areas_match_oracle(capi):
    cv2 = pytest.importorskip("cv2")
    from vsr_b200 import get_inpaint_area_by_mask

    rng = np.random.default_rng(4)
    for it in range(400):
        H, W = int(rng.integers(20, 200)), int(rng.integers(20, 300))
        if it % 2:
            m = ((rng.random((H, W)) < rng.choice([0.01, 0.03, 0.08])) * 255).astype(np.uint8)
            m = cv2.dilate(m, np.ones((3, 3), np.uint8), iterations=int(rng.integers(0, 3)))
        else:
            boxes = [(int(rng.integers(0, W)), int(rng.integers(0, W)), int(rng.integers(0, H)), int(rng.integers(0, H)))
                     for _ in range(int(rng.integers(1, 4)))]
            m = O.create_mask((H, W), [(min(a, b), max(a, b), min(c, d), max(c, d)) for a, b, c, d in boxes])
        h = max(1, int(W * 3 / 16))
        for mult in (1, 8):
            assert get_inpaint_area_by_mask(W, H, h, (m > 127).astype(np.uint8)[:, :, None], mult) == \
                O.get_inpaint_area_by_mask(W, H, h, (m > 127).astype(np.uint8), mult), (H, W, it, mult)
    assert get_inpaint_area_by_mask(64, 32, 12, np.zeros((32, 64), np.uint8)) == []


def test_default_1080p_strip(capi):
    from vsr_b200 import get_inpaint_area_by_mask

    m = O.default_mask(1080, 1920)
    assert int((m > 0).sum()) == 191100
    assert get_inpaint_area_by_mask(1920, 1080, 360, m) == [(720, 1080, 0, 1920)]


def test_batch_generator_and_schedule_match_oracle(capi):
    from vsr_b200 import batch_generator
    from vsr_b200.inpaint_tools import window_schedule

    for n in (0, 1, 7, 49, 50, 100, 299, 300, 500, 1200):
        for mb in (1, 3, 50, 70):
            got = [len(b) for b in batch_generator(list(range(n)), mb)]
            assert got == [b - a for a, b in O.batch_generator(n, mb)], (n, mb)
    for T in (1, 4, 5, 6, 11, 24, 46, 50, 77):
        for stride, ref in ((5, 10), (3, 7), (1, 1)):
            assert window_schedule(T, stride, ref) == [(a, b) for a, b in O.window_schedule(T, stride, ref)]


def test_ctypes_prototypes_match_the_header(capi):
    """Every ctypes prototype has as many arguments as the C declaration, pointers where the header has pointers and 64-bit integers
    where it has uint64_t / int64_t (a drift here is a crash or silent garbage on the GPU box, not a test failure on the CPU)."""
    import ctypes as C
    import re

    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "vsr_b200.h")).read(), flags=re.S)
    decl = dict(re.findall(r"\b(vsr_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text))
    checked = 0
    for name, (_, argtypes) in capi._PROTOS.items():
        params = [p.strip() for p in decl[name].split(",")]
        params = [] if params == ["void"] else params
        assert len(params) == len(argtypes), (name, params, argtypes)
        for p, t in zip(params, argtypes):
            is_ptr_c = "*" in p
            is_ptr_py = t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or getattr(t, "_type_", None) is not None and isinstance(t, type) and issubclass(t, C._Pointer)
            if is_ptr_c:
                assert is_ptr_py or t is C.c_uint64 and False, (name, p, t)
            elif re.search(r"\b(uint64_t|int64_t)\b", p):
                assert t in (C.c_uint64, C.c_int64), (name, p, t)
            elif re.search(r"\bfloat\b", p):
                assert t is C.c_float, (name, p, t)
            elif re.search(r"\b(int|int32_t)\b", p):
                assert t in (C.c_int, C.c_int32), (name, p, t)
            checked += 1
    assert checked > 600


def test_bench_input_generators_equal_the_oracle_generators():
    """bench.py's device legs take their frames and masks from vsr_b200.synthetic (nothing under oracle/ on the measured path); the goldens
    and the CPU baseline use the oracle's generators — same pixels."""
    from oracle import sttn_oracle as O
    from vsr_b200 import synthetic as S

    for n, H, W, seed in ((5, 128, 192, 23), (2, 1080, 1920, 0), (3, 720, 1280, 301)):
        assert all(np.array_equal(a, b) for a, b in zip(O.synthetic_clip(n, H, W, seed=seed), S.synthetic_clip(n, H, W, seed=seed)))
        assert np.array_equal(O.default_mask(H, W), S.default_mask(H, W))


def test_runtime_stand_in_has_the_interface_of_the_device_runtime():
    """The pipelines are developed against tests/fake_rt.py and shipped against the ctypes wrapper classes (`_DeviceRuntime` ... `_GenRuntime`):
    duck typing hides a method that is a property on one side, or takes other arguments, until the first run on a GPU.  Every public member of
    the stand-in must exist on the device wrapper with the same kind and the same parameters (names, order, defaults)."""
    import inspect

    from fake_rt import FakeRuntime
    from vsr_b200.propainter_generator import _GenRuntime

    only_stand_in = {"upload_ints", "launches", "bufs", "layers", "graphs", "rescaled", "enforce", "fp16"}
    problems = []
    for name in dir(FakeRuntime):
        if name.startswith("_") or name in only_stand_in:
            continue
        fake, real = inspect.getattr_static(FakeRuntime, name), inspect.getattr_static(_GenRuntime, name, None)
        if real is None:
            problems.append(f"{name}: missing on the device wrapper")
        elif isinstance(fake, property) != isinstance(real, property):
            problems.append(f"{name}: property on one side, method on the other")
        elif not isinstance(fake, property):
            fs, rs = inspect.signature(inspect.unwrap(fake)), inspect.signature(real)
            fp = [(p.name, p.default) for p in fs.parameters.values()]
            rp = [(p.name, p.default) for p in rs.parameters.values()]
            if fp != rp:
                problems.append(f"{name}: stand-in {fs} != device wrapper {rs}")
    assert not problems, "\n".join(problems)
