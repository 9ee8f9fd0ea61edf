"""GPU: each kernel family through the C ABI against a torch fp64 reference of the same op (operands
pre-rounded to fp16, which is what the tcgen05 path multiplies).  Tolerances are relative to max|ref|."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
F = torch.nn.functional
AUTO = [(80, 15), (32, 6), (10, 5), (5, 3)]


def _h(a):
    return a.astype(np.float16).astype(np.float32)


def _relerr(got, want):
    return float(np.abs(got.astype(np.float64) - want).max() / max(1e-9, np.abs(want).max()))


@pytest.fixture(scope="module")
def ops(capi):
    if capi.lib().vsr_device_count() < 1:
        pytest.fail("GPU tests need a B200 (sm_100) device")
    from vsr_b200 import ops as o

    return o


@pytest.mark.parametrize("T,H,W,Cin,Cout,k,dil,lrelu,res", [
    (1, 4, 32, 64, 64, 1, 1, False, False),
    (1, 8, 32, 256, 256, 1, 1, False, False),
    (2, 12, 40, 64, 128, 3, 1, True, False),
    (2, 30, 160, 256, 256, 3, 1, True, False),
    (2, 30, 160, 256, 256, 3, 2, True, True),     # feed_forward.conv.0 geometry + residual
    (2, 30, 160, 256, 768, 1, 1, False, False),   # fused Q/K/V projection
    (1, 27, 45, 128, 64, 3, 1, True, False),      # ragged tile edges
    (1, 60, 108, 256, 256, 3, 1, True, True),     # sttn-det feature map
])
def test_conv_igemm(ops, T, H, W, Cin, Cout, k, dil, lrelu, res):
    rng = np.random.default_rng(T * 1000 + Cin + Cout + k + dil)
    x = _h(rng.standard_normal((T, H, W, Cin), dtype=np.float32))
    w = _h(rng.standard_normal((Cout, Cin, k, k), dtype=np.float32) / np.sqrt(Cin * k * k))
    b = (rng.standard_normal(Cout) * 0.1).astype(np.float32)
    r = rng.standard_normal((T, H, W, Cout), dtype=np.float32) if res else None
    got = ops.conv2d(x, w, b, ksize=k, dilation=dil, lrelu=lrelu, residual=r)
    y = F.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2).double(), torch.from_numpy(w).double(), torch.from_numpy(b).double(),
                 padding=dil * (k // 2), dilation=dil)
    if lrelu:
        y = F.leaky_relu(y, 0.2)
    y = y.permute(0, 2, 3, 1).numpy()
    if res:
        y = y + r
    assert np.isfinite(got).all()
    assert _relerr(got, y) < 2e-3


def test_conv_stride2_space_to_depth(ops):
    rng = np.random.default_rng(9)
    x = _h(rng.standard_normal((2, 60, 320, 64), dtype=np.float32))
    w = _h(rng.standard_normal((128, 64, 3, 3), dtype=np.float32) / 24.0)
    b = (rng.standard_normal(128) * 0.1).astype(np.float32)
    got = ops.conv2d_s2(x, w, b, lrelu=True)
    y = F.leaky_relu(F.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2).double(), torch.from_numpy(w).double(),
                              torch.from_numpy(b).double(), stride=2, padding=1), 0.2).permute(0, 2, 3, 1).numpy()
    assert _relerr(got, y) < 2e-3


def _attention_ref(q, k, v, patches):
    from oracle import sttn_oracle as O

    T, H, W, C = q.shape
    dk = C // len(patches)
    qt, kt, vt = (torch.from_numpy(a).permute(0, 3, 1, 2).double() for a in (q, k, v))
    outs = []
    for i, (pw, ph) in enumerate(patches):
        sl = slice(i * dk, (i + 1) * dk)
        a, b, c = (O._split_tokens(z[:, sl], pw, ph) for z in (qt, kt, vt))
        p = torch.softmax(a @ b.t() / np.sqrt(a.shape[-1]), dim=-1)
        outs.append(O._merge_tokens(p @ c, T, dk, H, W, pw, ph))
    return torch.cat(outs, 1).permute(0, 2, 3, 1).numpy()


@pytest.mark.parametrize("T,H,W,C,patches,sharp", [
    (1, 4, 8, 64, [(2, 2)], 1.0),
    (2, 30, 160, 64, [(5, 3)], 1.0),
    (3, 30, 160, 64, [(32, 6)], 1.0),          # ow = 5 -> padded token slots
    (3, 30, 160, 64, [(80, 15)], 1.0),         # 6 tokens x 76800 dims: split-K
    (5, 30, 160, 256, AUTO, 1.0),
    (5, 30, 160, 256, AUTO, 4.0),              # peaked softmax
    (17, 30, 160, 128, [(10, 5), (5, 3)], 1.0), # 17-frame window: 5440-token rows (second softmax variant)
    (2, 60, 108, 256, [(108, 60), (36, 20), (18, 10), (9, 5)], 1.0),   # sttn-det geometry
])
def test_patch_attention(ops, T, H, W, C, patches, sharp):
    rng = np.random.default_rng(T + H + W + C)
    q, k = (_h(rng.standard_normal((T, H, W, C), dtype=np.float32) * sharp) for _ in range(2))
    v = _h(rng.standard_normal((T, H, W, C), dtype=np.float32))
    got = ops.patch_attention(q, k, v, patches)
    assert np.isfinite(got).all()
    assert _relerr(got, _attention_ref(q, k, v, patches)) < 4e-3


def test_upsample2x_align_corners(ops):
    rng = np.random.default_rng(0)
    x = _h(rng.standard_normal((2, 30, 160, 64), dtype=np.float32))
    got = ops.upsample2x(x)
    y = F.interpolate(torch.from_numpy(x).permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=True)
    assert _relerr(got, y.permute(0, 2, 3, 1).double().numpy()) < 2e-3


@pytest.mark.parametrize("sw,sh,dw,dh", [(1920, 360, 640, 120), (852, 159, 640, 120), (640, 120, 640, 120), (3840, 720, 640, 120)])
def test_resize_u8_bit_exact(ops, sw, sh, dw, dh):
    from oracle import sttn_oracle as O

    src = np.random.default_rng(sw).integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    assert np.array_equal(ops.resize_u8(src, dw, dh), O.cv2_resize_linear_u8(src, dw, dh))
