"""vsr_b200 — B200-native (sm_100a) engine for the subtitle-removal inpaint hot path of
YaoFANGUK/video-subtitle-remover, behind the reference's own plug-in interface.

Public surface (same names / call conventions as the reference back-ends):
    STTNInpaint(device, model_path)(frames, mask)                 backend/inpaint/sttn_auto_inpaint.py:28
    STTNAutoInpaint(device, model_path, video_path, ...)(...)     backend/inpaint/sttn_auto_inpaint.py:167
    STTNDetInpaint(device, model_path)(frames, mask)              backend/inpaint/sttn_det_inpaint.py:23
    LamaInpaint(device, model_path)(frames, mask) / .inpaint      backend/inpaint/lama_inpaint.py:11
    PropainterInpaint(device, model_dir, sub_video_length)(frames, mask)   backend/inpaint/propainter_inpaint.py:138
    SubtitleDetect(video_path, sub_areas).detect_subtitle(img)    backend/tools/subtitle_detect.py:16
    create_mask / get_inpaint_area_by_mask / batch_generator      backend/tools/inpaint_tools.py
    InpaintMode                                                   backend/tools/constant.py:4
All compute goes through the C-ABI library built from csrc/ (include/vsr_b200.h); importing the
compute classes without that library, or running them without a B200, raises.
"""
from .constant import InpaintMode  # noqa: F401
from .config import config  # noqa: F401
from .inpaint_tools import batch_generator, create_mask, get_inpaint_area_by_mask  # noqa: F401
from .sttn_auto_inpaint import STTNAutoInpaint, STTNInpaint  # noqa: F401
from .sttn_det_inpaint import STTNDetInpaint  # noqa: F401
from .subtitle_detect import SubtitleDetect  # noqa: F401
from .lama_inpaint import LamaInpaint  # noqa: F401
from .pipeline import propainter_mode_frames, video_inpaint_frames  # noqa: F401
from .pipeline_async import video_inpaint_frames_overlapped  # noqa: F401
from .propainter_inpaint import PropainterInpaint  # noqa: F401

__all__ = ["STTNInpaint", "STTNAutoInpaint", "STTNDetInpaint", "LamaInpaint", "PropainterInpaint", "SubtitleDetect", "video_inpaint_frames", "video_inpaint_frames_overlapped", "propainter_mode_frames", "InpaintMode", "config", "create_mask", "get_inpaint_area_by_mask",
           "batch_generator"]
