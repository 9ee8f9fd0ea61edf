"""CPU: the N>1 host logic (chunk sharding + max-over-ranks) with a real 2-process gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vsr_b200.distributed import (batches_for_rank, chunk_ranges, chunks_for_rank, detect_video_sharded, inpaint_clip_sharded, max_over_ranks,
                                  sampled_frames_for_rank)


def test_chunk_ranges_match_reference_loop():
    # sttn_auto_inpaint.py:240-245: rec_time = ceil(len / clip_gap), exact clip_gap chunks, last one shorter
    assert chunk_ranges(300, 50) == [(i * 50, i * 50 + 50) for i in range(6)]
    assert chunk_ranges(299, 50)[-1] == (250, 299)
    assert chunk_ranges(0, 50) == []
    assert chunk_ranges(7, 50) == [(0, 7)]
    with pytest.raises(ValueError):
        chunk_ranges(10, 0)


@pytest.mark.parametrize("n,gap,world", [(300, 50, 1), (300, 50, 2), (300, 50, 4), (300, 50, 8), (1200, 50, 8), (49, 50, 8)])
def test_sharding_is_a_partition(n, gap, world):
    seen = []
    for r in range(world):
        seen += chunks_for_rank(n, gap, r, world)
    seen.sort()
    assert [c for c, _ in seen] == list(range(len(chunk_ranges(n, gap))))
    assert [rg for _, rg in seen] == chunk_ranges(n, gap)


class _MarkEngine:
    """Stand-in for STTNInpaint: writes the rank into the frames it is given (in place)."""

    def __init__(self, rank):
        self.rank = rank

    def inpaint_inplace(self, frames, mask):
        for f in frames:
            f.fill_(self.rank + 1)


def _worker(rank, world, port, n, gap):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frames = [torch.zeros(4, dtype=torch.int32) for _ in range(n)]
        mine = inpaint_clip_sharded(_MarkEngine(rank), frames, None, gap, rank, world)
        owner = torch.stack(frames)[:, 0].clone()  # 0 = untouched here
        dist.all_reduce(owner, op=dist.ReduceOp.SUM)  # test-only collective: every frame painted exactly once
        expect = torch.tensor([(c % world) + 1 for c, (s, e) in enumerate(chunk_ranges(n, gap)) for _ in range(s, e)], dtype=torch.int32)
        assert torch.equal(owner, expect)
        assert all(c % world == rank for c, _ in mine)
        assert max_over_ranks(float(rank + 1)) == float(world)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 230, 50), nprocs=2, join=True)


class _FakeDetector:
    """Stand-in for SubtitleDetect: frame i (0-based) carries a subtitle box iff 10 <= i < 40, jittered by i % 3 pixels."""

    def detect_subtitle(self, frame):
        i = int(frame[0])
        return [(100 + i % 3, 400, 300, 330 + i % 2)] if 10 <= i < 40 else []


def _detect_worker(rank, world, port, n, step):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frames = [torch.tensor([i]) for i in range(n)]
        got = detect_video_sharded(_FakeDetector(), frames, step, rank, world)
        want = detect_video_sharded(_FakeDetector(), frames, step, 0, 1)      # the single-process plan
        assert got == want and min(got) == 13 and len(got) > 20               # first sampled hit: frame number 13 (0-based 12)
    finally:
        dist.destroy_process_group()


def test_detection_sharding_two_processes_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_detect_worker, args=(2, port, 60, 3), nprocs=2, join=True)


def test_sampled_frames_and_batches_partition():
    for n, step, world in ((100, 3, 4), (7, 2, 8), (60, 4, 2)):
        allf = sorted(f for r in range(world) for f in sampled_frames_for_rank(n, step, r, world))
        assert allf == list(range(1, n + 1, step))
    for n, mb, world in ((300, 46, 4), (50, 46, 2), (7, 46, 3)):
        spans = sorted(b for r in range(world) for b in batches_for_rank(n, mb, r, world))
        assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


# ---- ProPainter: one sub-video sharded over two ranks (RAFT clips + generator windows round-robin, two exchanges) -----------------
def _exchange_worker(rank, world, port):
    import numpy as np
    from vsr_b200.distributed import Shard

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = Shard(rank, world)
        mine = {i: np.full((i + 1, 3), i, np.float32 if i % 2 else np.uint8) for i in range(5) if sh.owns(i)}
        got = sh.exchange(mine)
        assert sorted(got) == list(range(5))
        for i, a in got.items():
            assert a.shape == (i + 1, 3) and a.dtype == (np.float32 if i % 2 else np.uint8) and (a == i).all()
        assert sh.exchange({}) == {}                                   # nothing to exchange is not a hang
    finally:
        dist.destroy_process_group()


def test_shard_exchange_two_processes_gloo():
    from vsr_b200.distributed import Shard
    from vsr_b200.propainter_inpaint import flow_clips

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_exchange_worker, args=(2, port), nprocs=2, join=True)
    assert Shard(0, 1).exchange({3: 1}) == {3: 1}
    with pytest.raises(ValueError):
        Shard(2, 2)
    # propainter_inpaint.py:209-236: clip length by width, clips after the first start one frame early
    assert flow_clips(7, 192) == [(0, 7)] and flow_clips(30, 640) == [(0, 12), (11, 24), (23, 30)]
    assert flow_clips(10, 1280) == [(0, 4), (3, 8), (7, 10)] and flow_clips(5, 1920) == [(0, 2), (1, 4), (3, 5)]
    for n, w in ((30, 640), (80, 1280), (9, 1920)):                  # every consecutive pair is covered exactly once
        pairs = [p for a, b in flow_clips(n, w) for p in range(a, b - 1)]
        assert pairs == list(range(n - 1))


_PP_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "weights", "propainter")


def _propainter_worker(rank, world, port, out_dir):
    import sys

    import numpy as np

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here, os.path.join(here, "..", "tools")]
    from fake_rt import FakeRuntime
    from make_golden_propainter import inputs
    from vsr_b200.distributed import Shard
    from vsr_b200.propainter_inpaint import PropainterInpaint

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(4)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frames, mask = inputs()[:2]                                   # 7 frames: windows at 0 and 5; RAFT clips of 4 -> (0,4) (3,7)
        eng = PropainterInpaint("cuda:0", _PP_DIR, runtime=FakeRuntime())
        eng.raft_clip = 4
        out = np.stack(eng.inpaint(frames, mask, Shard(rank, world)))
        np.save(os.path.join(out_dir, f"sharded_{rank}.npy"), out)
        if rank == 0:                                                 # the unsharded result of the same engine, for bit-exactness
            np.save(os.path.join(out_dir, "single.npy"), np.stack(eng.inpaint(frames, mask)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.slow
@pytest.mark.skipif(not all(os.path.exists(os.path.join(_PP_DIR, f)) for f in ("ProPainter.pth", "raft-things.pth", "recurrent_flow_completion.pth")),
                    reason="ProPainter weights not staged under weights/propainter")
def test_propainter_sub_video_sharded_over_two_ranks_equals_single_rank(tmp_path):
    import numpy as np

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_propainter_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    single = np.load(tmp_path / "single.npy")
    a, b = np.load(tmp_path / "sharded_0.npy"), np.load(tmp_path / "sharded_1.npy")
    assert np.array_equal(a, b) and np.array_equal(a, single)
    golden = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "propainter_real.npz"))["comp"]
    d = np.abs(a.astype(np.int32) - golden)
    assert d.max() <= 3 and (d > 0).mean() < 0.02


# ---- BASELINE config 4 sharded: detection + batches of every interval over two ranks == the single-process loop --------------------
class _SampledDetector(_FakeDetector):
    SAMPLE_STEP = 3

    def detect_subtitle(self, frame):
        i = int(frame[0, 0, 0]) * 256 + int(frame[0, 0, 1])
        return [(100, 400, 300, 330)] if 10 <= i < 130 else []

    def scan_frames(self, frames, sections=None, on_frame=None):
        return detect_video_sharded(self, frames, self.SAMPLE_STEP, 0, 1)


def _mark_model(batch, mask):
    """stand-in for STTNDetInpaint: output = input + 1 where the mask is set (pure function of its batch)"""
    import numpy as np

    return [np.where(mask[:, :, None] > 0, f + 1, f).astype(np.uint8) for f in batch]


def _config4_worker(rank, world, port, n):
    import numpy as np
    from vsr_b200.distributed import video_inpaint_frames_sharded
    from vsr_b200.pipeline import video_inpaint_frames

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frames = []
        for i in range(n):
            f = np.full((360, 640, 3), 7, np.uint8)
            f[0, 0, 0], f[0, 0, 1] = i // 256, i % 256
            frames.append(f)
        det = _SampledDetector()
        out, sub, se = video_inpaint_frames_sharded(frames, det, _mark_model, rank, world)
        want, wsub, wse = video_inpaint_frames(frames, det, _mark_model)
        assert sub == wsub and se == wse and len(se) >= 1 and len(out) == len(want) == n
        changed = torch.tensor([int(not np.array_equal(o, f)) for o, f in zip(out, frames)])
        equal_where_changed = all(np.array_equal(o, w) for o, w, c in zip(out, want, changed) if c)
        assert equal_where_changed                                     # this rank's batches carry the single-process result
        total = changed.clone()
        dist.all_reduce(total, op=dist.ReduceOp.SUM)                   # test-only collective: every inpainted frame done by exactly one rank
        expect = torch.tensor([int(not np.array_equal(w, f)) for w, f in zip(want, frames)])
        assert torch.equal(total, expect) and int(changed.sum()) > 0 and int(changed.sum()) < int(expect.sum())
    finally:
        dist.destroy_process_group()


def test_config4_sharded_two_processes_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_config4_worker, args=(2, port, 140), nprocs=2, join=True)
