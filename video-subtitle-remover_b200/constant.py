"""Mode enums, same members/values as backend/tools/constant.py:4-12 so `InpaintMode(value)` round-trips."""
from enum import Enum, unique


@unique
class InpaintMode(Enum):
    STTN_AUTO = "sttn-auto"
    STTN_DET = "sttn-det"
    LAMA = "lama"
    PROPAINTER = "propainter"
    OPENCV = "opencv"
