"""LAMA on the B200: drop-in for backend/inpaint/lama_inpaint.py `LamaInpaint`.

The reference loads the TorchScript `big-lama.pt` and calls `model(image, mask)` (lama_inpaint.py:13-28).  Here the
same FFCResNetGenerator (github.com/advimman/lama, saicinpainting/training/modules/ffc.py, as exported in that
file) is laid out as one CUDA graph per input size on the device-tensor runtime of the C ABI (`vsr_rt_*`,
include/vsr_b200.h):

* every convolution — 7x7 stem/head, the three stride-2 down convs, the 3x3 `l2l/g2l/l2g` convs and the 1x1 spectral
  convs of the 18 FFC residual blocks, the three transposed convs (as dense 3x3 convs over a zero-inserted input) —
  runs on the tcgen05 implicit-GEMM kernel with batch-norm, bias and ReLU folded into the epilogue;
* `padding_mode='reflect'` / `ReflectionPad2d`: one reflect-pad launch per FFC layer feeds all of its 3x3 convs, which
  run on the padded grid and store only the interior (crop in the conv epilogue);
* local (128 ch) and global (384 ch) features live in ONE 512-channel NHWC tensor, so `convl2l(x_l) + convg2l(x_g)` is a
  single conv over its channels and ConcatTupleLayer is free; the other convs read channel-slice views through TMA;
* FourierUnit: cuFFT R2C/C2R over the 192 interleaved channels (fp32 inside, `norm='ortho'`), the 1x1 conv on the
  (re, im)-interleaved spectrum in between on tensor cores;
* input packing (img/255*(1-m), m, symmetric padding to x8) and the output blend + u8 truncation are device kernels that
  restate lama_util.py:12-80 and lama_inpaint.py:25-27 operation by operation.

No CPU fallback: without the library or an sm_100 device the constructor raises `VsrError`.
"""
import ctypes as C
from typing import Dict, List, Sequence, Union

import numpy as np

from . import _capi
from .dbnet import _DeviceRuntime, _Tensor, _r
from .inpaint_tools import get_inpaint_area_by_mask

N_BLOCKS = 18
EPS = 1e-5


def load_lama_weights(path_or_dict) -> Dict[str, np.ndarray]:
    """TorchScript big-lama.pt (torch is only the container), an .npz of the same tensors, or a name -> array dict;
    names relative to `model.generator.model.`."""
    if isinstance(path_or_dict, dict):
        return {k: np.asarray(v, np.float32) for k, v in path_or_dict.items()}
    path = str(path_or_dict)
    if path.endswith(".npz"):
        z = np.load(path)
        return {k: z[k].astype(np.float32) for k in z.files}
    import torch

    sd = torch.jit.load(path, map_location="cpu").state_dict()
    pre = "model.generator.model."
    return {k[len(pre):]: v.float().numpy() for k, v in sd.items() if k.startswith(pre) and not k.endswith("num_batches_tracked")}


class _LamaRuntime(_DeviceRuntime):
    """The detector's runtime wrapper plus the LAMA-only entry points; tensors are [n, h, w, cp] batches."""

    def conv_ex(self, lid, x, y, relu, out_coff=0, crop=(0, 0)):
        _capi.check(self.L.vsr_rt_conv_ex(self.h, lid, x.ptr, x.n, x.h, x.w, y.ptr, y.cp, out_coff, relu, 1.0, 1.0, crop[0], crop[1], y.h, y.w))

    def pad(self, x, y, top, left, reflect=1):
        _capi.check(self.L.vsr_rt_pad(self.h, x.ptr, x.n, x.h, x.w, x.cp, y.ptr, y.h, y.w, top, left, reflect))

    def zero_upsample(self, x, y):
        _capi.check(self.L.vsr_rt_zero_upsample2x(self.h, x.ptr, x.n, x.h, x.w, x.cp, y.ptr))

    def add_slices(self, relu, a, b, y, channels):
        _capi.check(self.L.vsr_rt_add_slices(self.h, relu, a.ptr, a.cp, b.ptr, b.cp, y.ptr, y.cp, channels, y.pixels, 1.0, 1.0))

    def residual_add(self, x32, y, x, init):
        _capi.check(self.L.vsr_rt_residual_add(self.h, x32, y.ptr, x.ptr, x.pixels * x.cp, 1 if init else 0))

    def fft_r2c(self, x, y):
        _capi.check(self.L.vsr_rt_fft_r2c(self.h, x.ptr, x.n, x.h, x.w, x.c, x.cp, y.ptr))

    def fft_c2r(self, x, y):
        _capi.check(self.L.vsr_rt_fft_c2r(self.h, x.ptr, y.n, y.h, y.w, y.c, y.ptr, y.cp))

    def lama_input(self, img, mask, y, slot=0):
        _capi.check(self.L.vsr_rt_lama_input(self.h, _capi.ptr(img, C.c_uint8), _capi.ptr(mask, C.c_uint8), img.shape[0], img.shape[1],
                                             y.ptr + slot * y.h * y.w * y.cp * 2, y.h, y.w, y.cp, slot))

    def lama_output(self, pred, ih, iw, slot=0) -> np.ndarray:
        out = np.empty((ih, iw, 3), np.uint8)
        _capi.check(self.L.vsr_rt_lama_output(self.h, pred.ptr + slot * pred.h * pred.w * pred.cp * 2, pred.w, pred.cp, 1.0, ih, iw, slot,
                                              _capi.ptr(out, C.c_uint8)))
        return out


def _view(t: _Tensor, c0: int, c: int) -> _Tensor:
    """channels [c0, c0 + c) of an NHWC tensor: same pitch, pointer advanced by c0 fp16 elements."""
    return _Tensor(t.ptr + 2 * c0, c, t.h, t.w, t.cp, n=t.n)


def _bn_fold(w, p):
    s = w[f"{p}.weight"] / np.sqrt(w[f"{p}.running_var"] + np.float32(EPS))
    return s.astype(np.float32), (w[f"{p}.bias"] - w[f"{p}.running_mean"] * s).astype(np.float32)


class _Program:
    def __init__(self):
        self.steps, self.inp, self.out, self.graph = [], None, None, None


class LamaNetwork:
    """big-lama's forward(image, mask) on the device; one compiled program (CUDA graph) per padded input size."""

    def __init__(self, weights, device="cuda:0", runtime=None):
        self.w = load_lama_weights(weights)
        self._rt = runtime if runtime is not None else _LamaRuntime(device)
        self._programs: Dict[tuple, _Program] = {}
        self._layers: Dict[str, int] = {}

    def __del__(self):
        rt = getattr(self, "_rt", None)
        if rt is not None:
            try:
                rt.close()
            except Exception:
                pass
            self._rt = None

    # ---- layers (weights are uploaded once and shared by every resolution)
    def _conv(self, key, weight, bias, cin_pitch, stride=1, pad=0):
        lid = self._layers.get((key, cin_pitch))
        if lid is None:
            cout, cin, kh, kw = weight.shape
            lid = self._layers[(key, cin_pitch)] = self._rt.conv_create(weight, bias if bias is not None else np.zeros(cout, np.float32), cout,
                                                                       cin, cin_pitch, kh, kw, stride, pad, pad, 1, 1, False)
        return lid

    def _compile(self, N: int, H: int, W: int) -> _Program:
        if H % 8 or W % 8 or H < 16 or W < 16:
            raise _capi.VsrError("LAMA input must be padded to multiples of 8 (>= 16)")
        rt, w, prog = self._rt, self.w, _Program()
        run = prog.steps.append

        def new(c, h, wd) -> _Tensor:
            cp = _r(c, 64)
            return _Tensor(rt.alloc(N * h * wd * cp * 2), c, h, wd, cp, n=N)

        # model.0/1: ReflectionPad2d(3) + conv7x7 4->64 + BN + ReLU.  The 4 input channels sit in a 16-channel tensor.
        prog.inp = new(16, H, W)
        s, b = _bn_fold(w, "1.bn_l")
        w7 = np.zeros((64, 16, 7, 7), np.float32)
        w7[:, :4] = w["1.ffc.convl2l.weight"] * s[:, None, None, None]
        p0, stem = new(16, H + 6, W + 6), new(64, H, W)
        lid = self._conv("stem", w7, b, p0.cp, 1, 3)
        run(lambda: rt.pad(prog.inp, p0, 3, 3))
        run(lambda lid=lid: rt.conv_ex(lid, p0, stem, 1, 0, (3, 3)))

        # model.2-4: 3x3 stride-2 reflect convs.  P''[a] = x[reflect(a - 2)] ([h+4, w+4]); the stride-2 pad-1 conv of P''
        # is the wanted output shifted by one, so the epilogue keeps the window starting at (1, 1).
        x = stem
        for name, cout in (("2", 128), ("3", 256), ("4", 512)):
            if name == "4":   # local 128 | global 384 in one tensor
                sl, bl = _bn_fold(w, "4.bn_l")
                sg, bg = _bn_fold(w, "4.bn_g")
                wt = np.concatenate([w["4.ffc.convl2l.weight"] * sl[:, None, None, None], w["4.ffc.convl2g.weight"] * sg[:, None, None, None]])
                bias = np.concatenate([bl, bg])
            else:
                s, bias = _bn_fold(w, f"{name}.bn_l")
                wt = w[f"{name}.ffc.convl2l.weight"] * s[:, None, None, None]
            pp, y = new(x.c, x.h + 4, x.w + 4), new(cout, x.h // 2, x.w // 2)
            lid = self._conv(f"down{name}", wt, bias, pp.cp, 2, 1)
            run(lambda x=x, pp=pp: rt.pad(x, pp, 2, 2))
            run(lambda lid=lid, pp=pp, y=y: rt.conv_ex(lid, pp, y, 1, 0, (1, 1)))
            x = y

        # model.5-22: FFCResnetBlock x 18 on X = [local 128 | global 384]
        h, wd = x.h, x.w
        X = x
        P = new(512, h + 2, wd + 2)
        Y1, Y2 = new(512, h, wd), new(512, h, wd)
        G, G2 = new(384, h, wd), new(384, h, wd)
        S1, S2, S3 = new(192, h, wd), new(192, h, wd), new(192, h, wd)
        F1, F2 = new(384, h, wd // 2 + 1), new(384, h, wd // 2 + 1)

        def ffc(src: _Tensor, dst: _Tensor, p: str):
            sl, bl = _bn_fold(w, f"{p}.bn_l")
            sg, bg = _bn_fold(w, f"{p}.bn_g")
            wa = np.concatenate([w[f"{p}.ffc.convl2l.weight"], w[f"{p}.ffc.convg2l.weight"]], 1) * sl[:, None, None, None]
            la = self._conv(f"{p}.A", wa, bl, P.cp, 1, 1)
            lb = self._conv(f"{p}.B", w[f"{p}.ffc.convl2g.weight"] * sg[:, None, None, None], bg, P.cp, 1, 1)
            s1, b1 = _bn_fold(w, f"{p}.ffc.convg2g.conv1.1")
            l1 = self._conv(f"{p}.c1", w[f"{p}.ffc.convg2g.conv1.0.weight"] * s1[:, None, None, None], b1, src.cp)
            sf, bf = _bn_fold(w, f"{p}.ffc.convg2g.fu.bn")
            lf = self._conv(f"{p}.fu", w[f"{p}.ffc.convg2g.fu.conv_layer.weight"] * sf[:, None, None, None], bf, F1.cp)
            l2 = self._conv(f"{p}.c2", w[f"{p}.ffc.convg2g.conv2.weight"] * sg[:, None, None, None], None, S3.cp)
            run(lambda: rt.pad(src, P, 1, 1))
            run(lambda: rt.conv_ex(la, P, dst, 1, 0, (1, 1)))                       # relu(bn_l(l2l(x_l) + g2l(x_g))) -> dst[:, 0:128]
            run(lambda: rt.conv_ex(lb, _view(P, 0, 128), G, 0, 0, (1, 1)))          # bn_g scale/shift folded into l2g(x_l)
            run(lambda: rt.conv_ex(l1, _view(src, 128, 384), S1, 1))                # SpectralTransform.conv1 on x_g
            run(lambda: rt.fft_r2c(S1, F1))
            run(lambda: rt.conv_ex(lf, F1, F2, 1))
            run(lambda: rt.fft_c2r(F2, S2))
            run(lambda: rt.elementwise(0, S1, S2, S3))                              # x + fu(x)
            run(lambda: rt.conv_ex(l2, S3, G2, 0))
            run(lambda: rt.add_slices(1, G, G2, _view(dst, 128, 384), 384))         # relu(bn_g(l2g + g2g)) -> dst[:, 128:512]

        X32 = rt.alloc(X.pixels * X.cp * 4)    # fp32 master of the residual stream; X is its fp16 rounding for the convs
        for b in range(5, 5 + N_BLOCKS):
            ffc(X, Y1, f"{b}.conv1")
            ffc(Y1, Y2, f"{b}.conv2")
            run(lambda b=b: rt.residual_add(X32, Y2, X, b == 5))                    # x = id + x

        # model.24-32: ConvTranspose2d(3, s2, p1, op1) + BN + ReLU = dense 3x3 conv (flipped, transposed kernel) over the
        # zero-inserted input
        x = X
        for i in (24, 27, 30):
            s, b = _bn_fold(w, str(i + 1))
            wt = np.ascontiguousarray(np.transpose(w[f"{i}.weight"], (1, 0, 2, 3))[:, :, ::-1, ::-1]) * s[:, None, None, None]
            bias = w[f"{i}.bias"] * s + b
            z, y = new(x.c, 2 * x.h, 2 * x.w), new(wt.shape[0], 2 * x.h, 2 * x.w)
            lid = self._conv(f"up{i}", wt, bias, z.cp, 1, 1)
            run(lambda x=x, z=z: rt.zero_upsample(x, z))
            run(lambda lid=lid, z=z, y=y: rt.conv_ex(lid, z, y, 1))
            x = y

        # model.33-35: ReflectionPad2d(3) + conv7x7 64->3 (+ bias) + sigmoid; 3 output channels ride in an 8-channel conv
        wl = np.zeros((8, 64, 7, 7), np.float32)
        wl[:3] = w["34.weight"]
        bl = np.zeros(8, np.float32)
        bl[:3] = w["34.bias"]
        pl, logit, prog.out = new(64, H + 6, W + 6), new(8, H, W), new(8, H, W)
        lid = self._conv("head", wl, bl, pl.cp, 1, 3)
        run(lambda x=x: rt.pad(x, pl, 3, 3))
        run(lambda lid=lid: rt.conv_ex(lid, pl, logit, 0, 0, (3, 3)))
        run(lambda: rt.elementwise(3, logit, None, prog.out))

        for st in prog.steps:       # one eager pass creates the FFT plans and the stride-2 staging, then record the graph
            st()
        rt.overflow()
        rt.capture_begin()
        try:
            for st in prog.steps:
                st()
        finally:
            prog.graph = rt.capture_end()
        return prog

    MAX_BATCH = 4   # the reference's mini-batch (lama_inpaint.py:38); one CUDA graph per (batch, H, W)

    def forward_u8(self, image: np.ndarray, mask: np.ndarray) -> np.ndarray:
        """`LamaInpaint.inpaint` minus the model loading: image HxWx3 u8 (channel order untouched, lama_util.py:12-28),
        mask HxW u8 (> 0 = hole) -> HxWx3 u8."""
        return self.forward_batch([image], [mask])[0]

    def forward_batch(self, images: Sequence[np.ndarray], masks: Sequence[np.ndarray]) -> List[np.ndarray]:
        """Same-size images through the network, MAX_BATCH per graph launch (batch-norm is in eval mode: the result of an
        image does not depend on what it is batched with)."""
        imgs = [np.ascontiguousarray(i, np.uint8) for i in images]
        msks = [np.ascontiguousarray(m if m.ndim == 2 else m[:, :, 0], np.uint8) for m in masks]
        if len(imgs) != len(msks) or not imgs:
            raise ValueError("one mask per image")
        h, wd = msks[0].shape
        for i, m in zip(imgs, msks):
            if i.ndim != 3 or i.shape != (h, wd, 3) or m.shape != (h, wd):
                raise ValueError("expected images [H,W,3] and masks [H,W] of one common size")
        H, W = _r(h, 8), _r(wd, 8)
        if H - h >= h or W - wd >= wd:
            raise ValueError("image too small for symmetric padding to a multiple of 8")
        rt, out = self._rt, []
        for s0 in range(0, len(imgs), self.MAX_BATCH):
            n = min(self.MAX_BATCH, len(imgs) - s0)
            prog = self._programs.get((n, H, W))
            if prog is None:
                prog = self._programs[(n, H, W)] = self._compile(n, H, W)
            self._last = prog
            for j in range(n):
                rt.lama_input(imgs[s0 + j], msks[s0 + j], prog.inp, j)
            rt.graph_launch(prog.graph)
            out.extend(rt.lama_output(prog.out, h, wd, j) for j in range(n))
            if rt.overflow():
                raise _capi.VsrError("LAMA activations left the fp16 range on this frame")
        return out

    @property
    def launch_count(self) -> int:
        return self._rt.launch_count

    def time_network(self, iters: int = 10) -> float:
        """ms per replay of the network graph of the last-used (batch, size) (inputs already on the device)."""
        import time

        prog = self._last
        rt = self._rt
        rt.graph_launch(prog.graph)
        rt.sync()
        t0 = time.perf_counter()
        for _ in range(iters):
            rt.graph_launch(prog.graph)
        rt.sync()
        return (time.perf_counter() - t0) / iters * 1e3


class LamaInpaint:
    """backend/inpaint/lama_inpaint.py:11 — same constructor, `inpaint`, `_inpaint_batch` and `__call__`."""

    def __init__(self, device="cuda:0", model_path="big-lama.pt", runtime=None) -> None:
        self.device = device
        self.model = LamaNetwork(model_path, device, runtime)

    def inpaint(self, image: np.ndarray, mask: np.ndarray) -> np.ndarray:
        """lama_inpaint.py:17-28 (PIL inputs are converted like lama_util.get_image does)."""
        return self.model.forward_u8(np.array(image), np.array(mask))

    def _inpaint_batch(self, images: Sequence[np.ndarray], masks: Sequence[np.ndarray]) -> List[np.ndarray]:
        """lama_inpaint.py:30-66: mini-batches of 4 frames per network launch, like the reference."""
        return self.model.forward_batch([np.array(i) for i in images], [np.array(m) for m in masks])

    def __call__(self, input_frames: List[np.ndarray], input_mask: np.ndarray) -> List[np.ndarray]:
        """lama_inpaint.py:68-114: strips of height int(W*3/16) around the mask rows at native resolution; each strip is
        replaced whole by the network output; input frames are not modified."""
        mask = input_mask if input_mask.ndim == 2 else input_mask[:, :, 0]
        H, W = mask.shape[:2]
        areas = get_inpaint_area_by_mask(W, H, int(W * 3 / 16), mask)
        frames = [f.copy() for f in input_frames]
        # every strip is cropped from the untouched frames and only then are the results written back, in area order (:86-107):
        # strips may overlap
        comps = [self._inpaint_batch([f[y0:y1] for f in frames], [mask[y0:y1]] * len(frames)) for (y0, y1, _, _) in areas]
        for (y0, y1, _, _), strips in zip(areas, comps):
            for f, c in zip(frames, strips):
                f[y0:y1] = c
        return frames


__all__ = ["LamaInpaint", "LamaNetwork", "load_lama_weights"]
