"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference: from /root/reference in this container, from the byte-for-byte copy of its
Python modules under baseline/_ref (oracle/build_ref.py; git-ignored, travels with the gpurun snapshot) on the GPU box.

Used by tools/make_golden*.py and by the `-m "not gpu"` oracle-pinning tests to validate the restatements in oracle/ and to (re)generate
tests/golden/*.npz, and by bench.py's CPU legs (`--impl reference`, `cpu_baseline` kind "reference") to time the reference's own
`STTNInpaint.__call__` on the host cores.  The `-m gpu` tests and __graft_entry__.smoke() never import this.

Recipe follows SURVEY.md Appendix A.4: three stub modules (backend.config, matplotlib, fsplit) are
inserted before anything from `backend` is imported; weights are always passed as explicit paths
(never through ModelConfig / FFmpegCLI, which try to write into the read-only reference tree).
"""
import collections
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("VSR_REFERENCE_ROOT", "/root/reference")
_COPY = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")
if not os.path.isdir(os.path.join(REF_ROOT, "backend", "inpaint")) and os.path.isdir(os.path.join(_COPY, "backend", "inpaint")):
    REF_ROOT = _COPY


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "backend", "inpaint"))


def _value(v):
    return types.SimpleNamespace(value=v)


class _Cfg:  # defaults from backend/config.py:43-103
    sttnNeighborStride = _value(5)
    sttnReferenceLength = _value(10)
    sttnMaxLoadNum = _value(50)
    propainterMaxLoadNum = _value(70)
    subtitleAreaDeviationPixel = _value(10)
    subtitleYXAxisDifferencePixel = _value(10)
    subtitleAreaYAxisDifferencePixel = _value(20)
    subtitleAreaPixelToleranceYPixel = _value(20)
    subtitleAreaPixelToleranceXPixel = _value(20)
    subtitleTimelineBackwardFrameCount = _value(3)
    subtitleTimelineForwardFrameCount = _value(3)
    hardwareAcceleration = _value(False)

    def getSttnMaxLoadNum(self):
        return max(self.sttnMaxLoadNum.value, self.sttnNeighborStride.value * self.sttnReferenceLength.value)


_installed = False


def install():
    """Insert the stubs and put the reference on sys.path.  Idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    cfg = types.ModuleType("backend.config")
    cfg.config = _Cfg()
    cfg.tr = collections.defaultdict(lambda: collections.defaultdict(lambda: "{}"))
    cfg.BASE_DIR = os.path.join(REF_ROOT, "backend")
    cfg.VERSION = "1.4.0"
    sys.modules["backend.config"] = cfg
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        for s in ("patches", "path", "pyplot"):
            m = types.ModuleType("matplotlib." + s)
            setattr(mpl, s, m)
            sys.modules["matplotlib." + s] = m
        sys.modules["matplotlib.path"].Path = object
        sys.modules["matplotlib"] = mpl
    fs = types.ModuleType("fsplit")
    fsf = types.ModuleType("fsplit.filesplit")
    fsf.Filesplit = object
    fs.filesplit = fsf
    sys.modules["fsplit"] = fs
    sys.modules["fsplit.filesplit"] = fsf
    sys.path.insert(0, REF_ROOT)
    import backend  # noqa: E402

    backend.config = cfg
    _installed = True


def weights_path(kind: str) -> str:
    p = {
        "sttn-auto": "backend/models/sttn-auto/infer_model.pth",
        "sttn-det": "backend/models/sttn-det/sttn.pth",
    }[kind]
    return os.path.join(REF_ROOT, p)
