#!/usr/bin/env python
"""TEST INFRASTRUCTURE — recipe that puts the UNMODIFIED reference implementation of the hot path where it can travel to the GPU box.

The reference is Python: its "build" is a file copy.  From the sources where they lie under /root/reference this copies the Python
modules of `backend/inpaint`, `backend/tools` and `backend/scenedetect` (1.2 MB of .py files; no models, no ffmpeg binaries, no GUI)
into baseline/_ref/backend/ (the place the bench contract reserves for the unmodified reference), byte for byte.  baseline/_ref/ is listed in .gitignore (the copy never enters the history — reference sources
are not part of this repository) and not in .gpurunignore, so it travels with the snapshot like a compiled `_ref` binary would.

Users: `bench.py --impl reference` and the `cpu_baseline` leg (kind "reference": the reference's own `STTNInpaint.__call__`,
backend/inpaint/sttn_auto_inpaint.py:43-97, timed on the box's host cores), through oracle/ref_import.py which resolves the reference
root to /root/reference when that exists and to baseline/_ref otherwise.  Runs only where /root/reference exists (this container)."""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("VSR_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(os.path.dirname(HERE), "baseline", "_ref")
PACKAGES = ("inpaint", "tools", "scenedetect")


def main(quiet=False) -> bool:
    if not os.path.isdir(os.path.join(SRC, "backend", "inpaint")):
        if not quiet:
            print(f"[build_ref] {SRC} not present: nothing to do")
        return False
    n = 0
    pairs = [(os.path.join(SRC, "backend", "__init__.py"), os.path.join(DST, "backend", "__init__.py"))]
    for pkg in PACKAGES:
        for root, _, files in os.walk(os.path.join(SRC, "backend", pkg)):
            for f in files:
                if f.endswith(".py"):
                    s = os.path.join(root, f)
                    pairs.append((s, os.path.join(DST, os.path.relpath(s, SRC))))
    for s, d in pairs:
        if os.path.exists(d) and filecmp.cmp(s, d, shallow=False):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        n += 1
    if not quiet:
        print(f"[build_ref] {len(pairs)} reference modules under {DST} ({n} copied)")
    return True


if __name__ == "__main__":
    sys.exit(0 if main() else 1)
