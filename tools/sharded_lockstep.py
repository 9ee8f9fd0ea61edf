#!/usr/bin/env python
"""One chunk with its windows dealt over `world` ranks (vsr_sttn_shard_*), every rank its own engine on ONE GPU, driven in lock step from one
thread through the three phases; the two all-gathers are device-to-device copies between the engines' exchange buffers (under NCCL:
tools/run_sharded_check.py).  Run as a script so that the GPU test can isolate it in a child process:
    python tools/sharded_lockstep.py <world> <T>
Exit code 0 = the frames every rank hands back equal the unsharded call (world 1: bit for bit; else <= 2 grey levels on < 5 % of the pixels:
other windows share a launch, the split-K attention heads sum in another order)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(world: int, T: int) -> int:
    import torch
    from oracle import sttn_oracle as O
    from vsr_b200 import STTNInpaint, _capi
    from vsr_b200.sttn_auto_inpaint import _DevicePointer

    w = {k: v.numpy() for k, v in O.random_weights(0).items()}
    H, W = 270, 480
    frames = O.synthetic_clip(T, H, W, seed=77)
    mask = O.default_mask(H, W)
    want = STTNInpaint("cuda:0", w)(frames, mask)
    engines = [STTNInpaint("cuda:0", w) for _ in range(world)]

    def exchange(ptrs, nbytes):
        views = [torch.as_tensor(_DevicePointer(p, nbytes * world), device="cuda:0") for p in ptrs]
        for dst in range(world):
            for src in range(world):
                if src != dst:
                    views[dst][src * nbytes:(src + 1) * nbytes].copy_(views[src][src * nbytes:(src + 1) * nbytes])
        torch.cuda.synchronize()

    for attempt in range(2):
        outs = [[f.copy() for f in frames] for _ in range(world)]
        info = [engines[r].shard_begin(outs[r], mask, r, world) for r in range(world)]
        if world > 1:
            exchange([i[0] for i in info], info[0][1])
        for e in engines:
            e.shard_windows()
        if world > 1:
            exchange([i[2] for i in info], info[0][3])
        try:
            for r in range(world):
                engines[r].shard_finish(outs[r])
            break
        except _capi.VsrRangeError:          # every rank sees the same gathered flags: rank 0 raises first, all repeat on the exact path
            assert attempt == 0
            for e in engines:
                e.set_option("attn_direct", 0)
    worst, changed = 0, 0.0
    for rank in range(world):
        for f in range(T):
            if f % world != rank:
                assert np.array_equal(outs[rank][f], frames[f]), f"rank {rank} touched frame {f}"
            else:
                d = np.abs(outs[rank][f].astype(np.int32) - want[f])
                worst, changed = max(worst, int(d.max())), max(changed, float((d > 0).mean()))
    print(f"[lockstep] world={world} T={T}: max |diff| {worst}, largest changed fraction {changed:.2e}")
    if world == 1:
        return 0 if worst == 0 else 1
    return 0 if worst <= 2 and changed < 0.05 else 1


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]), int(sys.argv[2])))
