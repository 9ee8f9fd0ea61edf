"""Multi-GPU sharding of the hot path: one process per GPU, the independent unit is a
chunk of `clip_gap` frames (sttn_auto_inpaint.py:242-245 — every reference frame a window needs lies in
its own chunk), chunk c goes to rank c mod world.  No data-path collective; torch.distributed is used
only for the barrier and the max-over-ranks reduction of timings."""
from typing import List, Tuple


def chunk_ranges(n_frames: int, clip_gap: int) -> List[Tuple[int, int]]:
    """[(start, end)] exactly as the chunk loop of STTNAutoInpaint.__call__ (sttn_auto_inpaint.py:240-245)."""
    if clip_gap <= 0:
        raise ValueError("clip_gap must be positive")
    return [(s, min(s + clip_gap, n_frames)) for s in range(0, n_frames, clip_gap)]


def chunks_for_rank(n_frames: int, clip_gap: int, rank: int, world: int) -> List[Tuple[int, Tuple[int, int]]]:
    """(chunk index, (start, end)) owned by `rank`: round-robin, so consecutive chunks run concurrently."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return [(c, r) for c, r in enumerate(chunk_ranges(n_frames, clip_gap)) if c % world == rank]


def max_over_ranks(value: float, device=None) -> float:
    """Max of a scalar over all ranks (identity without an initialised process group)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def inpaint_clip_sharded(engine, frames, mask, clip_gap: int, rank: int, world: int):
    """Run this rank's chunks of an in-memory clip in place and return [(chunk index, (start, end))]."""
    mine = chunks_for_rank(len(frames), clip_gap, rank, world)
    for _, (s, e) in mine:
        engine.inpaint_inplace(frames[s:e], mask)
    return mine


# ---- detection + per-interval inpainting (BASELINE config 4 / LAMA): frames and batches are independent units ----------
def sampled_frames_for_rank(n_frames: int, step: int, rank: int, world: int) -> List[int]:
    """1-based numbers of the frames this rank runs the text detector on: the sampled frames 1, 1+step, ...
    (subtitle_detect.py:105) dealt round-robin, so every rank gets an even spread over the video."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    sampled = list(range(1, n_frames + 1, max(step, 1)))
    return sampled[rank::world]


def gather_detections(local: dict) -> dict:
    """Union of the per-rank {frame number: boxes} dictionaries on every rank.  This is the one real exchange step of the
    detection pass: gap filling and `unify_regions` (subtitle_detect.py:112-215) walk the whole dictionary in key order."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return dict(local)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local)
    merged = {}
    for p in parts:
        merged.update(p)
    return dict(sorted(merged.items()))


def detect_video_sharded(detector, frames, step: int, rank: int, world: int) -> dict:
    """`SubtitleDetect.scan_frames` over an in-memory clip with the sampled frames sharded across ranks: every rank
    detects its share, the hits are exchanged, and every rank finishes with the same planned dictionary."""
    from . import subtitle_plan as P

    local = {}
    for no in sampled_frames_for_rank(len(frames), step, rank, world):
        boxes = detector.detect_subtitle(frames[no - 1])
        if boxes:
            local[no] = boxes
    return P.drop_empty(P.unify_regions(P.gap_fill(gather_detections(local), step)))


def batches_for_rank(n_items: int, max_batch: int, rank: int, world: int) -> List[Tuple[int, int]]:
    """(start, end) of the `batch_generator` batches (inpaint_tools.py:7-29) of one interval that `rank` inpaints: sttn-det
    and LAMA treat every batch independently (main.py:323-326), so batches are dealt round-robin."""
    from .inpaint_tools import batch_generator

    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    out, s = [], 0
    for i, b in enumerate(batch_generator(list(range(n_items)), max_batch)):
        if i % world == rank:
            out.append((s, s + len(b)))
        s += len(b)
    return out


def video_inpaint_frames_sharded(frames, detector, model, rank: int, world: int):
    """BASELINE config 4 on `world` GPUs (the in-memory loop of SubtitleRemover.video_inpaint, main.py:260-333, like
    pipeline.video_inpaint_frames): every rank detects its share of the sampled frames, ONE all_gather_object makes the hit dictionary
    whole, every rank plans the same intervals and masks, and the `batch_generator` batches of all intervals are dealt round-robin —
    a batch is an independent unit for sttn-det / LAMA (main.py:323-326).  Returns (output frames, frame dictionary, interval map): the
    frames of this rank's batches are inpainted, all others are the inputs (each rank owns the output segments of its batches)."""
    from .config import config
    from .inpaint_tools import batch_generator, create_mask
    from .pipeline import interval_boxes, plan_intervals

    n = len(frames)
    sub_list = detect_video_sharded(detector, frames, detector.SAMPLE_STEP, rank, world)
    start_end = plan_intervals(sub_list, n)
    size = frames[0].shape[:2]
    out = list(frames)
    k = 0
    for s, e in sorted(start_end.items()):
        mask = None
        for batch in batch_generator(list(range(s - 1, e)), config.getSttnMaxLoadNum()):   # 0-based frame indices of the batch
            if len(batch) < 1:
                continue
            if k % world == rank:
                if mask is None:
                    mask = create_mask(size, interval_boxes(sub_list, s, e))
                out[batch[0]:batch[-1] + 1] = model([frames[i] for i in batch], mask)
            k += 1
    return out, sub_list, start_end


# ---- ProPainter (BASELINE config 5): one sub-video over several GPUs -------------------------------------------------------------
class Shard:
    """rank / world plus the one exchange the sharded ProPainter path needs (propainter_inpaint.PropainterInpaint.inpaint(shard=...)):
    units (RAFT clips, generator windows) are dealt round-robin, and `exchange` turns the per-rank {unit index: ndarray} dictionaries
    into the union on every rank.  Arrays travel as torch.distributed broadcasts from their owner — through device memory over NVLink
    with the NCCL backend, through host memory with gloo (the CPU tests); only their shapes go through the object channel."""

    def __init__(self, rank: int, world: int):
        if not 0 <= rank < world:
            raise ValueError("rank out of range")
        self.rank, self.world = rank, world

    def owns(self, unit: int) -> bool:
        return unit % self.world == self.rank

    def exchange(self, mine: dict) -> dict:
        import numpy as np
        import torch
        import torch.distributed as dist

        if self.world == 1:
            return dict(mine)
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("Shard.exchange needs an initialised torch.distributed process group")
        metas = [None] * self.world
        dist.all_gather_object(metas, [(k, tuple(v.shape), v.dtype.str) for k, v in sorted(mine.items())])
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        out = {}
        for src, meta in enumerate(metas):
            for k, shape, dt in meta:
                host = np.ascontiguousarray(mine[k]) if src == self.rank else np.empty(shape, np.dtype(dt))
                t = torch.from_numpy(host).to(device)
                dist.broadcast(t, src=src)
                out[k] = t.cpu().numpy() if src != self.rank or device.type != "cpu" else host
        return out
