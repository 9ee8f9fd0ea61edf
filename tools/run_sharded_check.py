#!/usr/bin/env python
"""Under torchrun (one process per GPU, NCCL): one 1080p chunk with its windows dealt over the ranks (`STTNInpaint.inpaint_chunk_sharded`,
NCCL all-gathers of the reference-frame features and of the window predictions, device to device) against the unsharded
call computed by the same rank on its own GPU — the frames a rank hands back must equal it (<= 2 grey levels: split-K summation order).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/run_sharded_check.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from vsr_b200 import STTNInpaint
    from vsr_b200 import synthetic as S

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p = os.path.join(ROOT, "weights", "sttn-auto", "infer_model.pth")
    eng = STTNInpaint(torch.device("cuda", local), p if os.path.exists(p) else S.random_sttn_weights(0))
    ok = True
    for T, H, W in ((50, 1080, 1920), (23, 270, 480)):
        frames = S.synthetic_clip(T, H, W, seed=7)
        mask = S.default_mask(H, W)
        want = eng(frames, mask)
        work = [f.copy() for f in frames]
        mine = eng.inpaint_chunk_sharded(work, mask, rank, world)
        dmax = max(int(np.abs(work[f].astype(np.int32) - want[f]).max()) for f in mine)
        same = dmax <= (0 if world == 1 else 2)      # other windows share a launch: split-K summation order
        untouched = all(np.array_equal(work[f], frames[f]) for f in range(T) if f not in mine)
        t0 = time.perf_counter()
        for _ in range(3):
            eng.inpaint_chunk_sharded([f.copy() for f in frames], mask, rank, world)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        flag = torch.tensor([int(same and untouched)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok &= bool(flag.item())
        if rank == 0:
            print(f"[sharded-check] world={world} {T}x{W}x{H}: own frames equal to the unsharded call on every rank (max |diff| here {dmax}): {bool(flag.item())}; {dt * 1e3:.1f} ms per chunk "
                  f"({T / dt:.0f} frames/s)", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
