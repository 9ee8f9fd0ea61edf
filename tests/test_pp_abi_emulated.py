"""The ProPainter operators (csrc/pp_ops.cuh + their `vsr_rt_*` entry points in csrc/engine.cu) have not run on a B200 yet.
Until they do, this file runs their REAL SOURCE on the CPU, through the product's own call chain:

    vsr_b200 wrapper method (_GenRuntime.unfold7s3 ...)  ->  ctypes with the prototypes of vsr_b200/_capi.py
      ->  the entry point cut out of csrc/engine.cu (argument checks, scratch buffers, grid sizes; tests/emu/make_abi_emu.py rewrites only
          the `<<<...>>>` launch syntax)  ->  the kernel of csrc/pp_ops.cuh compiled by g++ over tests/emu/cuda_emu.h (__half with
          round-to-nearest-even, uint4, blockIdx/threadIdx; warp shuffles and __syncthreads run one OS thread per CUDA thread in lockstep)

and compares every operator with the numpy stand-in of the runtime (tests/fake_rt.py) that the CPU parity tests of the whole pipeline are
built on.  So the stand-in is no longer only a transcription: index arithmetic, layouts, reductions, launch grids, argument order and the
ctypes prototypes of the shipped code are executed here.  What this cannot show: anything about the device itself (memory model, launch
limits, tensor-core convolutions, performance): tests/test_gpu_zz_pp_ops.py runs the same cases (tests/pp_op_cases.py) on the device, and the
gated stage tests of tests/test_gpu_raft.py follow it in the bring-up plan."""
import ctypes as C
import os
import subprocess
import sys

import pytest

from conftest import ROOT
from pp_op_cases import CASES, Dual, HostBackend

EMU = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "video-subtitle-remover_b200", "csrc")


@pytest.fixture(scope="module")
def backend():
    from vsr_b200 import _capi

    build = os.path.join(EMU, "build")
    out, gen = os.path.join(build, "libabi_emu.so"), os.path.join(build, "abi_emu.cpp")
    srcs = [os.path.join(EMU, f) for f in ("make_abi_emu.py", "abi_prelude.h", "cuda_emu.h")] + [os.path.join(CSRC, "engine.cu"), os.path.join(CSRC, "pp_ops.cuh"),
                                                                                                os.path.join(ROOT, "include", "vsr_b200.h")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(s) for s in srcs):
        os.makedirs(build, exist_ok=True)
        subprocess.run([sys.executable, srcs[0], os.path.join(CSRC, "engine.cu"), gen], check=True)
        subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-fsanitize=alignment", "-fno-sanitize-recover=alignment", "-I", os.path.join(EMU, "stubs"), "-I", EMU, gen, "-o", out], check=True)
    L = C.CDLL(out)
    for name, (res, args) in _capi._PROTOS.items():          # the shipped prototypes, applied to the host build of the same entry points
        if hasattr(L, name):
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
    L.emu_rt_create.restype = C.c_void_p
    L.emu_launches.restype, L.emu_launches.argtypes = C.c_long, [C.c_void_p]
    saved, _capi._lib = _capi._lib, L                          # _capi.check() reads the error text from the library in use
    yield HostBackend(L)
    _capi._lib = saved


@pytest.mark.parametrize("name,case,params", CASES, ids=[c[0] for c in CASES])
def test_operator(backend, name, case, params):
    case(lambda seed: Dual(backend, seed), *params)
