"""CPU: the whole ProPainter pipeline with the REAL kernels in it.  `HybridRuntime` (tests/hybrid_rt.py) is the fp16-storage stand-in of the
runtime whose ProPainter operators execute the real source of csrc/pp_ops.cuh through the product's wrappers and the real `vsr_rt_*` entry
points (host build, tests/emu/), at the pipeline's own shapes, pitches, channel-slice views and index lists; convolutions stay numpy, and so do the kernels that
need lockstep warps (one OS thread per CUDA thread in the host build): by default instance norm, layer norm and the window attention, in the
opt-in variant only the window attention (all three are covered one by one in tests/test_pp_abi_emulated.py).  Output against the frames
of the UNMODIFIED reference (fp32): >= 50 dB in the hole (last runs 59.2 dB / max 6 grey levels with 27 kernels, 60.7 dB / max 5 with 29),
bit-exact outside it."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
DIR = os.path.join(ROOT, "weights", "propainter")
pytestmark = pytest.mark.skipif(not all(os.path.exists(os.path.join(DIR, f)) for f in ("ProPainter.pth", "raft-things.pth", "recurrent_flow_completion.pth")),
                                reason="ProPainter weights not staged under weights/propainter")


@pytest.mark.slow
@pytest.mark.parametrize("on_numpy", [("instnorm", "layernorm", "window_attention"),
                                      pytest.param(("window_attention",), marks=pytest.mark.skipif(os.environ.get("VSR_SLOW_TESTS") != "1", reason="4.5 minutes: set "
                                                   "VSR_SLOW_TESTS=1 (last run: 60.7 dB in the hole, max 5 grey levels)"))],
                         ids=["27_kernels", "29_kernels"])
def test_pipeline_with_real_kernels_reproduces_reference_frames(on_numpy):
    from hybrid_rt import _SIG, HybridRuntime
    from make_golden_propainter import inputs
    from oracle import propainter_oracle as P
    from oracle import sttn_oracle as O
    from pp_op_cases import load_emu_library
    from vsr_b200 import _capi
    from vsr_b200.propainter_inpaint import PropainterInpaint

    lib = load_emu_library()
    saved, _capi._lib = _capi._lib, lib
    try:
        z = np.load(os.path.join(GOLDEN, "propainter_real.npz"))
        frames, mask = inputs()[:2]
        rt = HybridRuntime(lib, on_numpy=on_numpy)
        out = np.stack(PropainterInpaint("cuda:0", DIR, runtime=rt).inpaint(frames, mask))
    finally:
        _capi._lib = saved
    assert set(rt.real_calls) == set(_SIG) - set(on_numpy), sorted(set(_SIG) - set(on_numpy) - set(rt.real_calls))     # every kernel really ran
    assert not rt._flag                                                                                         # no fp16 overflow anywhere
    hole = np.stack(P.read_mask(mask, len(frames))[1]) > 0
    assert np.array_equal(out[~hole], np.stack(frames)[~hole])
    d = np.abs(out.astype(np.int32) - z["comp"])
    psnr = O.psnr_u8(out[hole].astype(np.float32), z["comp"][hole].astype(np.float32))
    assert psnr >= 50 and d.max() <= 12, (psnr, int(d.max()))
