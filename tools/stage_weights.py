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
         "V5/ch_det/inference.yml": "backend/models/V5/ch_det/inference.yml"}


def main(quiet=False):
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
