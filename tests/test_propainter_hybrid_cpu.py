"""CPU: the whole ProPainter pipeline with the REAL kernels in it.  `HybridRuntime` (tests/hybrid_rt.py) is the fp16-storage stand-in of the
runtime whose ProPainter operators — all 30 kernels of csrc/pp_ops.cuh — execute from their real source through the product's wrappers and
the real `vsr_rt_*` entry points (host build, tests/emu/), at the pipeline's own shapes, pitches, channel-slice views and index lists; only
the tensor-core convolutions (and the correlation GEMM) stay numpy, with fp16 operands and fp16 storage.  Output against the frames of the
UNMODIFIED reference (fp32): >= 50 dB in the hole (last run 60.7 dB, max 5 grey levels on 5e-5 of the pixels), bit-exact outside it."""
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
def test_pipeline_with_real_kernels_reproduces_reference_frames():
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
        rt = HybridRuntime(lib)
        out = np.stack(PropainterInpaint("cuda:0", DIR, runtime=rt).inpaint(frames, mask))
    finally:
        _capi._lib = saved
    assert set(rt.real_calls) == set(_SIG), sorted(set(_SIG) - set(rt.real_calls))     # every kernel really ran
    assert not rt._flag                                                                                         # no fp16 overflow anywhere
    hole = np.stack(P.read_mask(mask, len(frames))[1]) > 0
    assert np.array_equal(out[~hole], np.stack(frames)[~hole])
    d = np.abs(out.astype(np.int32) - z["comp"])
    psnr = O.psnr_u8(out[hole].astype(np.float32), z["comp"][hole].astype(np.float32))
    assert psnr >= 50 and d.max() <= 12, (psnr, int(d.max()))
