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
import pytest

from pp_op_cases import CASES, Dual, HostBackend, load_emu_library


@pytest.fixture(scope="module")
def backend():
    from vsr_b200 import _capi

    L = load_emu_library()
    saved, _capi._lib = _capi._lib, L                          # _capi.check() reads the error text from the library in use
    yield HostBackend(L)
    _capi._lib = saved


@pytest.mark.parametrize("name,case,params", CASES, ids=[c[0] for c in CASES])
def test_operator(backend, name, case, params):
    case(lambda seed: Dual(backend, seed), *params)


@pytest.mark.parametrize("name,case,params", CASES, ids=[c[0] for c in CASES])
def test_device_flavour_of_the_suite_rehearsed_on_the_host_build(backend, monkeypatch, name, case, params):
    """tests/test_gpu_zz_pp_ops.py drives the cases through `DeviceBackend` (vsr_rt_create / alloc / upload / launch / download); the host build
    restates those memory entry points with malloc / memcpy, so that this code path — not only the zero-copy one above — runs on the CPU too."""
    from pp_op_cases import DeviceBackend
    from vsr_b200 import dbnet

    monkeypatch.setattr(dbnet, "_device_index", lambda device: 0)
    made = []

    def make(seed):
        made.append(Dual(DeviceBackend(), seed))
        return made[-1]

    try:
        case(make, *params)
    finally:
        for d in made:
            d.close()


def test_copy_bytes_through_the_wrapper(backend):
    import numpy as np

    from pp_op_cases import bind_wrapper

    rt = bind_wrapper(backend.lib)
    src, dst = np.arange(64, dtype=np.float32), np.zeros(64, np.float32)
    rt.copy_bytes(src.ctypes.data + 32, dst.ctypes.data + 64, 96)                 # (src, dst, bytes): 24 floats from src[8:] to dst[16:]
    assert np.array_equal(dst[16:40], src[8:32]) and not dst[:16].any() and not dst[40:].any()
