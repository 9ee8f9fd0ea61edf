"""GPU: every operator of the ProPainter path (csrc/pp_ops.cuh behind its `vsr_rt_*` entry points) against the numpy stand-in of the runtime,
one operator per case — the cases of tests/pp_op_cases.py, which the CPU suite runs on a host build of the same source
(tests/test_pp_abi_emulated.py).  This is the first rung of the ProPainter bring-up: it needs no weights and localises a
device-side fault to one kernel before the stage tests of tests/test_gpu_raft.py run.  The file sorts last on purpose: these kernels had not
run on a B200 when the file was written, and a device fault here must not take the validated suites down with it.
Tolerances are those of the CPU run: exact for data movement and decisions, 1-4 fp16 ulp (+1e-3 absolute where sums cancel) for arithmetic."""
import pytest

from pp_op_cases import CASES, DeviceBackend, Dual

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,case,params", CASES, ids=[c[0] for c in CASES])
def test_operator_on_device(capi, name, case, params):
    made = []

    def make(seed):
        made.append(Dual(DeviceBackend(), seed))
        return made[-1]

    try:
        case(make, *params)
    finally:
        for d in made:
            d.close()
