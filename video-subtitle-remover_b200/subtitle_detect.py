"""Host-side mirror of backend/tools/subtitle_detect.py `SubtitleDetect`: the text
detector runs on the B200 (vsr_b200.dbnet.TextDetector instead of paddleocr's CPU TextDetection), the sampling /
gap-fill / unify / interval logic is vsr_b200.subtitle_plan (bit-exact restatements)."""
import os
from typing import Dict, Iterable, List

from . import subtitle_plan as P
from .dbnet import TextDetector


class SubtitleDetect:
    """Drop-in for backend/tools/subtitle_detect.py:16.  `model_dir` is what ModelConfig.DET_MODEL_DIR points at
    (backend/models/V5/ch_det for PP-OCRv5_server_det, backend/tools/model_config.py:17-23)."""

    SAMPLE_STEP = 3

    def __init__(self, video_path, sub_areas=None, model_dir=None, device="cuda:0"):
        self.video_path = video_path
        self.sub_areas = sub_areas if sub_areas is not None else []
        self.model_dir = model_dir or os.environ.get("VSR_DET_MODEL_DIR", "")
        self.device = device
        self._detector = None
        self._init_sample_step()

    def _init_sample_step(self):
        """subtitle_detect.py:29-39 (needs the video's fps; keeps the class default when there is no video)."""
        if not self.video_path:
            return
        import cv2

        cap = cv2.VideoCapture(self.video_path)
        fps = cap.get(cv2.CAP_PROP_FPS)
        cap.release()
        self.SAMPLE_STEP = P.sample_step_for_fps(fps)

    @property
    def text_detector(self) -> TextDetector:
        if self._detector is None:
            self._detector = TextDetector(self.model_dir, self.device)
        return self._detector

    def detect_subtitle(self, img) -> List[P.Box]:
        """subtitle_detect.py:56-82: boxes (xmin, xmax, ymin, ymax) of the text lines inside the selected areas."""
        out: List[P.Box] = []
        for res in self.text_detector.predict(img):
            polys = res["dt_polys"]
            if polys is None or len(polys) == 0:
                continue
            out.extend(P.filter_boxes(P.get_coordinates(polys.tolist()), self.sub_areas))
        return out

    def scan_frames(self, frames: Iterable, sections=None, on_frame=None) -> Dict[int, List[P.Box]]:
        """The loop body of subtitle_detect.py:96-132 over any frame iterator: detect every SAMPLE_STEP-th frame
        (1-based keys), fill gaps of at most 2*SAMPLE_STEP between hits, unify near-identical boxes, drop empty entries."""
        from .sttn_auto_inpaint import _in_ab_sections

        sampled, no = {}, 0
        for frame in frames:
            no += 1
            if not _in_ab_sections(no - 1, sections):
                continue
            if P.is_sampled(no, self.SAMPLE_STEP):
                boxes = self.detect_subtitle(frame)
                if boxes:
                    sampled[no] = boxes
            if on_frame is not None:
                on_frame(no)
        return P.drop_empty(P.unify_regions(P.gap_fill(sampled, self.SAMPLE_STEP)))

    @staticmethod
    def get_scene_div_frame_no(v_path, device="cuda:0"):
        """subtitle_detect.py:158-170 (frame numbers at which a new scene starts; ProPainter mode splits its intervals there, main.py:165):
        the ContentDetector scores on the B200 (vsr_b200.scene_detect), threshold / scene-list logic bit-exact on the host."""
        from .scene_detect import get_scene_div_frame_no

        return get_scene_div_frame_no(v_path, device)

    def find_subtitle_frame_no(self, sub_remover=None) -> Dict[int, List[P.Box]]:
        """subtitle_detect.py:84-132 on `self.video_path`."""
        import cv2

        cap = cv2.VideoCapture(self.video_path)
        total = cap.get(cv2.CAP_PROP_FRAME_COUNT)

        def frames():
            while cap.isOpened():
                ok, frame = cap.read()
                if not ok:
                    break
                yield frame

        def progress(no):
            if sub_remover is not None and total:
                sub_remover.progress_total = (100 * float(no) / float(total)) // 2

        try:
            return self.scan_frames(frames(), getattr(sub_remover, "ab_sections", None), progress)
        finally:
            cap.release()

    # the static interval helpers keep their reference names
    find_continuous_ranges = staticmethod(P.find_continuous_ranges)
    find_continuous_ranges_with_same_mask = staticmethod(P.find_continuous_ranges_with_same_mask)
    filter_and_merge_intervals = staticmethod(P.filter_and_merge_intervals)
    split_range_by_scene = staticmethod(P.split_range_by_scene)
    unify_regions = staticmethod(P.unify_regions)
