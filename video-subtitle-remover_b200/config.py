"""Hot-path knobs with the reference's names and defaults (backend/config.py:43-103).

The reference reads `config.<name>.value` at construction time of each back-end
(sttn_auto_inpaint.py:40-41,195); the engine does the same against this object.  When the host
application has already imported the reference's own `backend.config`, `adopt(backend.config.config)`
makes the engine follow the user's settings instead of the defaults.
"""
import types


def _v(x):
    return types.SimpleNamespace(value=x)


class _Config:
    def __init__(self):
        self.sttnNeighborStride = _v(5)          # backend/config.py:89
        self.sttnReferenceLength = _v(10)        # :91
        self.sttnMaxLoadNum = _v(50)             # :93
        self.subtitleAreaDeviationPixel = _v(10)  # :61
        self.subtitleSelectionAreas = _v("0.88,0.99,0.15,0.85")  # :43
        self.subtitleAreaPixelToleranceYPixel = _v(20)  # :65
        self.subtitleAreaPixelToleranceXPixel = _v(20)  # :66
        self.subtitleTimelineBackwardFrameCount = _v(3)  # :67
        self.subtitleTimelineForwardFrameCount = _v(3)   # :68
        self.subtitleYXAxisDifferencePixel = _v(10)      # :59
        self.propainterMaxLoadNum = _v(70)               # :100

    def getSttnMaxLoadNum(self):                 # :94
        return max(self.sttnMaxLoadNum.value, self.sttnNeighborStride.value * self.sttnReferenceLength.value)

    def adopt(self, other):
        for k in ("sttnNeighborStride", "sttnReferenceLength", "sttnMaxLoadNum", "subtitleAreaDeviationPixel",
                  "subtitleAreaPixelToleranceYPixel", "subtitleAreaPixelToleranceXPixel", "subtitleTimelineBackwardFrameCount",
                  "subtitleTimelineForwardFrameCount", "subtitleYXAxisDifferencePixel", "propainterMaxLoadNum"):
            if hasattr(other, k):
                getattr(self, k).value = getattr(other, k).value


config = _Config()
