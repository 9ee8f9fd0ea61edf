#!/usr/bin/env python
"""Copy the reference checkpoints the hot path needs into weights/ (git-ignored, travels with gpurun).
Only works where /root/reference exists (the build container); a no-op elsewhere."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VSR_REFERENCE_ROOT", "/root/reference")
FILES = {"sttn-auto/infer_model.pth": "backend/models/sttn-auto/infer_model.pth",
         "sttn-det/sttn.pth": "backend/models/sttn-det/sttn.pth",
         "V5/ch_det/inference.json": "backend/models/V5/ch_det/inference.json",
         "V5/ch_det/inference.pdiparams": "backend/models/V5/ch_det/inference.pdiparams",
         "V5/ch_det/inference.yml": "backend/models/V5/ch_det/inference.yml",
         "V5/ch_det_fast/inference.json": "backend/models/V5/ch_det_fast/inference.json",
         "V5/ch_det_fast/inference.pdiparams": "backend/models/V5/ch_det_fast/inference.pdiparams",
         "V5/ch_det_fast/inference.yml": "backend/models/V5/ch_det_fast/inference.yml"}


def stage_lama(quiet=False):
    """big-lama: concatenate the 5 parts in fs_manifest.csv order into the TorchScript file (kept out of gpurun snapshots by
    .gpurunignore) and export its generator tensors as an fp16 .npz (102 MB) that travels to the GPU box."""
    src = os.path.join(REF, "backend", "models", "big-lama")
    dst = os.path.join(ROOT, "weights", "big-lama")
    pt, npz = os.path.join(dst, "big-lama.pt"), os.path.join(dst, "big-lama.npz")
    if not os.path.isdir(src):
        if not quiet:
            print(f"skip big-lama: {src} not found")
        return
    os.makedirs(dst, exist_ok=True)
    if not os.path.exists(pt):
        import csv

        with open(os.path.join(src, "fs_manifest.csv")) as f:
            parts = [r["filename"] for r in csv.DictReader(f)]
        with open(pt, "wb") as out:
            for part in parts:
                with open(os.path.join(src, part), "rb") as f:
                    shutil.copyfileobj(f, out)
        if not quiet:
            print(f"staged {pt}")
    if not os.path.exists(npz):
        import numpy as np
        import torch

        sd = torch.jit.load(pt, map_location="cpu").state_dict()
        pre = "model.generator.model."
        np.savez(npz, **{k[len(pre):]: v.numpy().astype(np.float16 if v.dim() == 4 else np.float32) for k, v in sd.items()
                         if k.startswith(pre) and not k.endswith("num_batches_tracked")})   # conv kernels fp16, batch-norm vectors fp32
        if not quiet:
            print(f"staged {npz}")


def stage_propainter_compact(quiet=False):
    """ProPainter.pth (158 MB fp32) -> ProPainter.f16.pth (79 MB): the same state dict with every tensor of two or more dimensions stored
    in fp16 (the device path rounds them to fp16 for the tensor cores anyway), vectors fp32.  The fp32 file is listed in .gpurunignore
    (snapshot cap 512 MiB); loaders fall back to the compact file when the fp32 one is absent."""
    d = os.path.join(ROOT, "weights", "propainter")
    src, dst = os.path.join(d, "ProPainter.pth"), os.path.join(d, "ProPainter.f16.pth")
    if not os.path.exists(src) or (os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src)):
        return
    import torch

    sd = torch.load(src, map_location="cpu")
    torch.save({k: (v.half() if v.is_floating_point() and v.dim() >= 2 else v) for k, v in sd.items()}, dst)
    if not quiet:
        print(f"staged {dst}")


def main(quiet=False):
    if "--lama" in sys.argv:
        stage_lama(quiet)
    stage_propainter_compact(quiet)
    for dst, src in FILES.items():
        s = os.path.join(REF, src)
        d = os.path.join(ROOT, "weights", dst)
        if not os.path.exists(s):
            if not quiet:
                print(f"skip {dst}: {s} not found")
            continue
        if os.path.exists(d) and os.path.getsize(d) == os.path.getsize(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        if not quiet:
            print(f"staged {d}")


if __name__ == "__main__":
    main()
