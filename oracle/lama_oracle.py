"""CPU oracle for the LAMA path — TEST INFRASTRUCTURE ONLY.

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; the product (vsr_b200.lama_inpaint)
never touches it.  It restates, in plain torch fp32 on the CPU:

  L1  get_image / pad_img_to_modulo / prepare_img_and_mask     backend/inpaint/utils/lama_util.py:12-80
  L2  the TorchScript `big-lama.pt` forward(image, mask)       backend/models/big-lama/ (5 parts, fs_manifest.csv);
      FFCResNetGenerator of github.com/advimman/lama (saicinpainting/training/modules/ffc.py), as exported:
      checked op by op against the script module's own `.code` and numerically against the module itself
  L3  LamaInpaint.inpaint / _inpaint_batch / __call__          backend/inpaint/lama_inpaint.py:17-114

Parity: PINNED — tests/test_lama_oracle.py compares `forward` with the unmodified TorchScript module on the real
weights (when weights/big-lama/big-lama.pt is staged) and with tests/golden/lama_real.npz (outputs of the unmodified
reference `LamaInpaint` produced by tools/make_golden.py).
"""
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

N_BLOCKS = 18          # model.5 .. model.22
EPS = 1e-5             # BatchNorm2d default


# ------------------------------------------------------------------------------------------------ weights
def weight_shapes() -> Dict[str, tuple]:
    """state_dict of `model.generator.model` without the num_batches_tracked counters (names relative to it)."""
    s: Dict[str, tuple] = {}

    def bn(prefix, c):
        for k in ("weight", "bias", "running_mean", "running_var"):
            s[f"{prefix}.{k}"] = (c,)

    s["1.ffc.convl2l.weight"] = (64, 4, 7, 7)
    bn("1.bn_l", 64)
    s["2.ffc.convl2l.weight"] = (128, 64, 3, 3)
    bn("2.bn_l", 128)
    s["3.ffc.convl2l.weight"] = (256, 128, 3, 3)
    bn("3.bn_l", 256)
    s["4.ffc.convl2l.weight"] = (128, 256, 3, 3)
    s["4.ffc.convl2g.weight"] = (384, 256, 3, 3)
    bn("4.bn_l", 128)
    bn("4.bn_g", 384)
    for b in range(5, 5 + N_BLOCKS):
        for c in ("conv1", "conv2"):
            p = f"{b}.{c}"
            s[f"{p}.ffc.convl2l.weight"] = (128, 128, 3, 3)
            s[f"{p}.ffc.convl2g.weight"] = (384, 128, 3, 3)
            s[f"{p}.ffc.convg2l.weight"] = (128, 384, 3, 3)
            s[f"{p}.ffc.convg2g.conv1.0.weight"] = (192, 384, 1, 1)
            bn(f"{p}.ffc.convg2g.conv1.1", 192)
            s[f"{p}.ffc.convg2g.fu.conv_layer.weight"] = (384, 384, 1, 1)
            bn(f"{p}.ffc.convg2g.fu.bn", 384)
            s[f"{p}.ffc.convg2g.conv2.weight"] = (384, 192, 1, 1)
            bn(f"{p}.bn_l", 128)
            bn(f"{p}.bn_g", 384)
    for i, (ci, co) in ((24, (512, 256)), (27, (256, 128)), (30, (128, 64))):
        s[f"{i}.weight"] = (ci, co, 3, 3)   # ConvTranspose2d layout [Cin, Cout, kh, kw]
        s[f"{i}.bias"] = (co,)
        bn(str(i + 1), co)
    s["34.weight"] = (3, 64, 7, 7)
    s["34.bias"] = (3,)
    return s


def load_weights(path: str) -> Dict[str, torch.Tensor]:
    """TorchScript big-lama.pt (torch.jit.load, CPU) or an .npz written by `save_weights` -> name -> fp32 tensor."""
    if path.endswith(".npz"):
        z = np.load(path)
        return {k: torch.from_numpy(z[k].astype(np.float32)) for k in z.files}
    sd = torch.jit.load(path, map_location="cpu").state_dict()
    pre = "model.generator.model."
    w = {k[len(pre):]: v.float() for k, v in sd.items() if k.startswith(pre) and not k.endswith("num_batches_tracked")}
    want = weight_shapes()
    assert set(w) == set(want) and all(tuple(w[k].shape) == want[k] for k in want), "unexpected big-lama state_dict"
    return w


def random_weights(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded weights of the big-lama architecture with activations that stay O(1) through the 18 residual blocks
    (conv gain 1/sqrt(fan_in), batch-norm statistics near identity) — for GPU parity runs without the 206 MB file."""
    g = torch.Generator().manual_seed(seed)
    w = {}
    for k, shp in weight_shapes().items():
        if k.endswith("running_var"):
            w[k] = 0.5 + torch.rand(shp, generator=g)
        elif k.endswith("running_mean"):
            w[k] = 0.1 * torch.randn(shp, generator=g)
        elif len(shp) == 1 and k.endswith(".weight"):          # BN gamma
            w[k] = 0.6 + 0.3 * torch.rand(shp, generator=g)
        elif len(shp) == 1:                                      # BN beta / conv bias
            w[k] = 0.05 * torch.randn(shp, generator=g)
        else:
            fan_in = shp[1] * shp[2] * shp[3] if not k.split(".")[0] in ("24", "27", "30") else shp[0] * shp[2] * shp[3] / 4
            w[k] = torch.randn(shp, generator=g) / float(np.sqrt(fan_in))
    return w


# ------------------------------------------------------------------------------------------------ L2: the network
def _bn(w, p, x):
    return F.batch_norm(x, w[f"{p}.running_mean"], w[f"{p}.running_var"], w[f"{p}.weight"], w[f"{p}.bias"], False, 0.0, EPS)


def _conv_reflect(x, weight, stride=1):
    """FFC's nn.Conv2d(..., padding=k//2, padding_mode='reflect', bias=False)."""
    p = weight.shape[-1] // 2
    if p:
        x = F.pad(x, (p, p, p, p), mode="reflect")
    return F.conv2d(x, weight, None, stride)


def fourier_unit(w, p, x):
    """FourierUnit.forward (ffc.py; as exported): rfftn ortho -> (re, im) interleaved on channels -> 1x1 conv + BN +
    ReLU -> complex -> irfftn ortho."""
    b, c, h, wd = x.shape
    f = torch.fft.rfftn(x, dim=(-2, -1), norm="ortho")
    f = torch.stack((f.real, f.imag), dim=-1).permute(0, 1, 4, 2, 3).contiguous().view(b, 2 * c, h, wd // 2 + 1)
    f = F.relu(_bn(w, f"{p}.bn", F.conv2d(f, w[f"{p}.conv_layer.weight"])))
    f = f.view(b, c, 2, h, wd // 2 + 1).permute(0, 1, 3, 4, 2).contiguous()
    return torch.fft.irfftn(torch.complex(f[..., 0], f[..., 1]), s=(h, wd), dim=(-2, -1), norm="ortho")


def spectral_transform(w, p, x):
    """SpectralTransform.forward with stride 1, enable_lfu=False: conv2(conv1(x) + fu(conv1(x)))."""
    y = F.relu(_bn(w, f"{p}.conv1.1", F.conv2d(x, w[f"{p}.conv1.0.weight"])))
    return F.conv2d(y + fourier_unit(w, f"{p}.fu", y), w[f"{p}.conv2.weight"])


def ffc_bn_act(w, p, xl, xg, stride=1):
    """FFC_BN_ACT with ReLU on both branches; xg is None when ratio_gin == 0; returns (local, global-or-None)."""
    ol = _conv_reflect(xl, w[f"{p}.ffc.convl2l.weight"], stride)
    if xg is not None:
        ol = ol + _conv_reflect(xg, w[f"{p}.ffc.convg2l.weight"], stride)
    ol = F.relu(_bn(w, f"{p}.bn_l", ol))
    if f"{p}.ffc.convl2g.weight" not in w:
        return ol, None
    og = _conv_reflect(xl, w[f"{p}.ffc.convl2g.weight"], stride)
    if xg is not None:
        og = og + spectral_transform(w, f"{p}.ffc.convg2g", xg)
    return ol, F.relu(_bn(w, f"{p}.bn_g", og))


def generator(w, x: torch.Tensor, taps: Dict[str, torch.Tensor] | None = None) -> torch.Tensor:
    """FFCResNetGenerator (model.0 .. model.35) on [B,4,H,W], H and W multiples of 8."""
    def tap(name, t):
        if taps is not None:
            taps[name] = t

    x = F.pad(x, (3, 3, 3, 3), mode="reflect")                                     # 0
    x = F.relu(_bn(w, "1.bn_l", F.conv2d(x, w["1.ffc.convl2l.weight"])))          # 1 (padding 0)
    tap("stem", x)
    x, _ = ffc_bn_act(w, "2", x, None, 2)
    x, _ = ffc_bn_act(w, "3", x, None, 2)
    xl, xg = ffc_bn_act(w, "4", x, None, 2)
    tap("down_l", xl)
    tap("down_g", xg)
    for b in range(5, 5 + N_BLOCKS):                                               # FFCResnetBlock
        yl, yg = ffc_bn_act(w, f"{b}.conv1", xl, xg)
        yl, yg = ffc_bn_act(w, f"{b}.conv2", yl, yg)
        xl, xg = xl + yl, xg + yg
        tap(f"block{b}_l", xl)
        tap(f"block{b}_g", xg)
    x = torch.cat((xl, xg), 1)                                                     # 23 ConcatTupleLayer
    for i in (24, 27, 30):
        x = F.conv_transpose2d(x, w[f"{i}.weight"], w[f"{i}.bias"], stride=2, padding=1, output_padding=1)
        x = F.relu(_bn(w, str(i + 1), x))
        tap(f"up{i}", x)
    x = F.pad(x, (3, 3, 3, 3), mode="reflect")                                     # 33
    return torch.sigmoid(F.conv2d(x, w["34.weight"], w["34.bias"]))               # 34, 35


def forward(w, image: torch.Tensor, mask: torch.Tensor, taps=None) -> torch.Tensor:
    """big-lama.pt forward(image [B,3,H,W] in 0..1, mask [B,1,H,W] in {0,1}) (the script module's own code)."""
    with torch.no_grad():
        mask = mask.to(image.dtype)
        masked = image * (1 - mask)
        pred = generator(w, torch.cat([masked, mask], 1), taps)
        return mask * pred + (1 - mask) * image


# ------------------------------------------------------------------------------------------------ L1 / L3: host side
def get_image(img: np.ndarray) -> np.ndarray:
    """lama_util.py:12-30: HWC / HW uint8 -> CHW float32 / 255."""
    a = img.copy()
    a = np.transpose(a, (2, 0, 1)) if a.ndim == 3 else a[np.newaxis, ...]
    return a.astype(np.float32) / 255


def pad_img_to_modulo(img: np.ndarray, mod: int = 8) -> np.ndarray:
    """lama_util.py:55-63: symmetric padding at the bottom / right up to a multiple of `mod`."""
    _, h, wd = img.shape
    oh, ow = -(-h // mod) * mod, -(-wd // mod) * mod
    return np.pad(img, ((0, 0), (0, oh - h), (0, ow - wd)), mode="symmetric")


def inpaint(w, image: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """LamaInpaint.inpaint (lama_inpaint.py:17-28): one image HxWx3 u8, mask HxW[,1] u8 -> HxWx3 u8."""
    h, wd = image.shape[:2]
    img = torch.from_numpy(pad_img_to_modulo(get_image(image)))[None]
    m = torch.from_numpy(pad_img_to_modulo(get_image(mask)))[None]
    out = forward(w, img, (m > 0) * 1)[0].permute(1, 2, 0).numpy()
    return np.clip(out * 255, 0, 255).astype("uint8")[:h, :wd]


def inpaint_batch(w, images: Sequence[np.ndarray], masks: Sequence[np.ndarray]) -> List[np.ndarray]:
    """LamaInpaint._inpaint_batch (lama_inpaint.py:30-66): mini-batches of 4 (batch-norm in eval mode: per-frame results
    do not depend on the batching)."""
    if len(images) == 1:
        return [inpaint(w, images[0], masks[0])]
    h, wd = images[0].shape[:2]
    out: List[np.ndarray] = []
    for s in range(0, len(images), 4):
        img = torch.from_numpy(np.stack([pad_img_to_modulo(get_image(i)) for i in images[s:s + 4]]))
        m = torch.from_numpy(np.stack([pad_img_to_modulo(get_image(k)) for k in masks[s:s + 4]]))
        res = forward(w, img, (m > 0) * 1).permute(0, 2, 3, 1).numpy()
        res = np.clip(res * 255, 0, 255).astype("uint8")
        out.extend(r[:h, :wd] for r in res)
    return out


def lama_call(w, input_frames: Sequence[np.ndarray], input_mask: np.ndarray) -> List[np.ndarray]:
    """LamaInpaint.__call__ (lama_inpaint.py:68-114): strips of height int(W*3/16) around the mask at native
    resolution, every strip replaced whole by the network output."""
    from oracle import sttn_oracle as O

    mask = input_mask[:, :, None]
    H, W = mask.shape[:2]
    areas = O.get_inpaint_area_by_mask(W, H, int(W * 3 / 16), mask)
    frames = [f.copy() for f in input_frames]
    for (y0, y1, _, _) in areas:
        comps = inpaint_batch(w, [f[y0:y1] for f in frames], [mask[y0:y1] for _ in frames])
        for f, c in zip(frames, comps):
            f[y0:y1] = c
    return frames
