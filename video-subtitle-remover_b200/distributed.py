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
