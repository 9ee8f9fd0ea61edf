#!/usr/bin/env python
"""GPU bring-up diagnostics: every kernel family in isolation against a torch fp32 reference, one
subprocess per case (a trapped kernel poisons its CUDA context, so cases must not share a process).

    python tools/gpu_diag.py            # run all cases, write gpurun_out/diag.log
    python tools/gpu_diag.py <case>     # run one case in this process
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def _h(a):
    return a.astype(np.float16).astype(np.float32)


def report(name, got, want, tol):
    d = np.abs(got.astype(np.float64) - want.astype(np.float64))
    scale = max(1e-6, float(np.abs(want).max()))
    bad = d > tol * scale
    info = dict(case=name, max_err=float(d.max()), mean_err=float(d.mean()), scale=scale, frac_bad=float(bad.mean()),
                finite=bool(np.isfinite(got).all()), ok=bool(not bad.any() and np.isfinite(got).all()))
    if bad.any():
        idx = np.argwhere(bad)[:6]
        info["first_bad"] = [(tuple(int(v) for v in i), float(got[tuple(i)]), float(want[tuple(i)])) for i in idx]
        # which trailing-dim slices are wrong (helps to spot a broken tap / channel group / row group)
        for ax in range(got.ndim):
            other = tuple(a for a in range(got.ndim) if a != ax)
            prof = bad.mean(axis=other)
            info[f"bad_profile_axis{ax}"] = [round(float(v), 3) for v in prof[:64]]
    print(json.dumps(info))
    return info["ok"]


def conv_case(name, T, H, W, Cin, Cout, k, dil, lrelu=False, residual=False, seed=0):
    import torch
    import torch.nn.functional as F
    from vsr_b200 import ops

    rng = np.random.default_rng(seed)
    x = _h(rng.standard_normal((T, H, W, Cin), dtype=np.float32))
    w = _h(rng.standard_normal((Cout, Cin, k, k), dtype=np.float32) / np.sqrt(Cin * k * k))
    b = rng.standard_normal(Cout).astype(np.float32) * 0.1
    res = rng.standard_normal((T, H, W, Cout), dtype=np.float32) if residual else None
    got = ops.conv2d(x, w, b, ksize=k, dilation=dil, lrelu=lrelu, residual=res)
    xt = torch.from_numpy(x).permute(0, 3, 1, 2).double()
    y = F.conv2d(xt, torch.from_numpy(w).double(), torch.from_numpy(b).double(), padding=dil * (k // 2), dilation=dil)
    if lrelu:
        y = F.leaky_relu(y, 0.2)
    y = y.permute(0, 2, 3, 1).numpy()
    if residual:
        y = y + res
    return report(name, got, y.astype(np.float32), 2e-3)


def conv_s2_case(name, T, H, W, Cin, Cout, seed=0):
    import torch
    import torch.nn.functional as F
    from vsr_b200 import ops

    rng = np.random.default_rng(seed)
    x = _h(rng.standard_normal((T, H, W, Cin), dtype=np.float32))
    w = _h(rng.standard_normal((Cout, Cin, 3, 3), dtype=np.float32) / np.sqrt(Cin * 9))
    b = rng.standard_normal(Cout).astype(np.float32) * 0.1
    got = ops.conv2d_s2(x, w, b, lrelu=True)
    y = F.leaky_relu(F.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2).double(), torch.from_numpy(w).double(),
                              torch.from_numpy(b).double(), stride=2, padding=1), 0.2)
    return report(name, got, y.permute(0, 2, 3, 1).numpy().astype(np.float32), 2e-3)


def attn_case(name, T, H, W, C, patches, seed=0, sharp=1.0):
    import torch
    from oracle import sttn_oracle as O
    from vsr_b200 import ops

    rng = np.random.default_rng(seed)
    q, k, v = (_h(rng.standard_normal((T, H, W, C), dtype=np.float32) * s) for s in (sharp, sharp, 1.0))
    got = ops.patch_attention(q, k, v, patches)
    dk = C // len(patches)
    qt, kt, vt = (torch.from_numpy(a).permute(0, 3, 1, 2).double() for a in (q, k, v))
    outs = []
    for i, (pw, ph) in enumerate(patches):
        sl = slice(i * dk, (i + 1) * dk)
        a, b, c = (O._split_tokens(z[:, sl], pw, ph) for z in (qt, kt, vt))
        p = torch.softmax(a @ b.t() / np.sqrt(a.shape[-1]), dim=-1)
        outs.append(O._merge_tokens(p @ c, T, dk, H, W, pw, ph))
    y = torch.cat(outs, 1).permute(0, 2, 3, 1).numpy().astype(np.float32)
    return report(name, got, y, 4e-3)


def upsample_case(name):
    import torch
    import torch.nn.functional as F
    from vsr_b200 import ops

    rng = np.random.default_rng(0)
    x = _h(rng.standard_normal((2, 30, 160, 64), dtype=np.float32))
    got = ops.upsample2x(x)
    y = F.interpolate(torch.from_numpy(x).permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=True)
    return report(name, got, y.permute(0, 2, 3, 1).numpy(), 2e-3)


def resize_case(name):
    from oracle import sttn_oracle as O
    from vsr_b200 import ops

    rng = np.random.default_rng(0)
    ok = True
    for (sw, sh, dw, dh) in [(1920, 360, 640, 120), (852, 159, 640, 120), (640, 120, 640, 120)]:
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        got = ops.resize_u8(src, dw, dh)
        ok &= report(f"{name}:{sw}x{sh}", got.astype(np.float32), O.cv2_resize_linear_u8(src, dw, dh).astype(np.float32), 0.0)
    return ok


def strip_case(name, T, real=False):
    from oracle import sttn_oracle as O
    from vsr_b200 import STTNInpaint

    if real:
        p = os.path.join(ROOT, "weights", "sttn-auto", "infer_model.pth")
        w = O.load_weights(p)
        eng = STTNInpaint("cuda:0", p)
    else:
        w = O.random_weights(0)
        eng = STTNInpaint("cuda:0", {k: v.numpy() for k, v in w.items()})
    strip = [O.cv2_resize_linear_u8(f, 640, 120) for f in O.synthetic_clip(T, 360, 1920, seed=21)]
    t0 = time.time()
    got = eng.inpaint([s.copy() for s in strip])
    t1 = time.time()
    taps = {}
    want = O.inpaint_strip(w, strip, taps=taps)
    enc = taps["encoder"].permute(0, 2, 3, 1).numpy()
    report(name + ":encoder(feats32)", eng.debug_read("feats32", enc.shape), enc, 4e-3)
    g = np.stack([c.astype(np.float32) for c in got])
    wv = np.stack([c.astype(np.float32) for c in want])
    print(json.dumps(dict(case=name, gpu_s=t1 - t0, psnr=O.psnr_u8(g, wv), launches=eng.launch_count,
                          dtype_match=[str(a.dtype) == str(b.dtype) for a, b in zip(got, want)])))
    return report(name, g, wv, 4.0 / 255.0)


def call_case(name, T, H, W):
    from oracle import sttn_oracle as O
    from vsr_b200 import STTNInpaint

    w = O.random_weights(0)
    eng = STTNInpaint("cuda:0", {k: v.numpy() for k, v in w.items()})
    frames = O.synthetic_clip(T, H, W, seed=5)
    mask = O.default_mask(H, W)
    got = eng(frames, mask)
    want = O.sttn_call(w, frames, mask)
    return report(name, np.stack(got).astype(np.float32), np.stack(want).astype(np.float32), 4.0 / 255.0)


AUTO = [(80, 15), (32, 6), (10, 5), (5, 3)]
CASES = {
    "resize": lambda n: resize_case(n),
    "upsample": lambda n: upsample_case(n),
    "conv1x1_c64_tiny": lambda n: conv_case(n, 1, 4, 32, 64, 64, 1, 1),
    "conv1x1_c256": lambda n: conv_case(n, 1, 8, 32, 256, 256, 1, 1),
    "conv3x3_c64": lambda n: conv_case(n, 1, 8, 32, 64, 64, 3, 1),
    "conv3x3_c64_to128": lambda n: conv_case(n, 2, 12, 40, 64, 128, 3, 1, lrelu=True),
    "conv3x3_c256_feat": lambda n: conv_case(n, 2, 30, 160, 256, 256, 3, 1, lrelu=True),
    "conv3x3_dil2_res": lambda n: conv_case(n, 2, 30, 160, 256, 256, 3, 2, lrelu=True, residual=True),
    "conv1x1_qkv768": lambda n: conv_case(n, 2, 30, 160, 256, 768, 1, 1),
    "conv3x3_odd": lambda n: conv_case(n, 1, 27, 45, 128, 64, 3, 1, lrelu=True),
    "conv_s2": lambda n: conv_s2_case(n, 2, 60, 320, 64, 128),
    "attn_1head_tiny": lambda n: attn_case(n, 1, 4, 8, 64, [(2, 2)]),
    "attn_1head_5x3": lambda n: attn_case(n, 2, 30, 160, 64, [(5, 3)]),
    "attn_1head_32x6": lambda n: attn_case(n, 3, 30, 160, 64, [(32, 6)]),
    "attn_1head_80x15": lambda n: attn_case(n, 3, 30, 160, 64, [(80, 15)]),
    "attn_4head_T5": lambda n: attn_case(n, 5, 30, 160, 256, AUTO),
    "attn_4head_T5_sharp": lambda n: attn_case(n, 5, 30, 160, 256, AUTO, sharp=4.0),
    "strip_rand_T3": lambda n: strip_case(n, 3),
    "strip_rand_T12": lambda n: strip_case(n, 12),
    "call_rand": lambda n: call_case(n, 6, 270, 480),
    "strip_real_T7": lambda n: strip_case(n, 7, real=True),
}


def main():
    only = None
    if len(sys.argv) > 2 and sys.argv[1] == "--only":
        only = sys.argv[2].split(",")
    elif len(sys.argv) > 1:
        name = sys.argv[1]
        ok = CASES[name](name)
        sys.exit(0 if ok else 1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "diag.log"), "w")
    summary = {}
    for name in CASES:
        if only and not any(name.startswith(o) for o in only):
            continue
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True, timeout=240)
            status = "ok" if r.returncode == 0 else f"FAIL({r.returncode})"
            out = r.stdout + ("\n[stderr]\n" + r.stderr[-3000:] if r.returncode != 0 else "")
        except subprocess.TimeoutExpired as e:
            status, out = "TIMEOUT", (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else str(e.stdout)
        summary[name] = status
        log.write(f"=== {name}: {status} ({time.time() - t0:.1f}s)\n{out}\n")
        log.flush()
        print(f"{name}: {status} ({time.time() - t0:.1f}s)", flush=True)
    log.write("SUMMARY " + json.dumps(summary) + "\n")
    log.close()


if __name__ == "__main__":
    main()
