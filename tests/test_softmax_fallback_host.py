"""CPU: the host side of the single-pass-softmax range guard (VSR_ERR_RANGE -> engine option attn_direct = 0 -> repeat), with a stand-in for
the C library: the synchronous call repeats once; in the two-deep chunk pipeline the chunk that raised is repeated from its untouched input
frames, and a chunk that was already in flight (it shares the cleared device flag) is collected into scratch copies and repeated too."""
import ctypes as C

import numpy as np
import pytest

from vsr_b200 import _capi
from vsr_b200.sttn_auto_inpaint import STTNInpaint


class FakeLib:
    """counts calls; `fail_collect` / `fail_frames` = tickets / call numbers that answer VSR_ERR_RANGE once"""

    def __init__(self):
        self.options, self.calls, self.tickets = [], [], 0
        self.fail_collect, self.fail_frames = set(), set()
        self.n_frames_calls = 0

    def vsr_last_error(self):
        return b"attention logits beyond the range of the single-pass softmax"

    def vsr_sttn_set_option(self, h, name, value):
        self.options.append((name.decode(), value))
        return 0

    def _paint(self, ptr_array, n, value):
        for i in range(n):
            buf = (C.c_uint8 * 12).from_address(ptr_array[i])
            for j in range(12):
                buf[j] = (buf[j] + value) % 256

    def vsr_sttn_inpaint_frames(self, h, pin, n, H, W, mask, pout):
        self.n_frames_calls += 1
        self.calls.append(("frames", self.n_frames_calls))
        if self.n_frames_calls in self.fail_frames:
            self.fail_frames.discard(self.n_frames_calls)
            return -5
        for i in range(n):      # out = in + 1 (reads the INPUT pointers: a repeat on already-painted frames would show)
            src = (C.c_uint8 * 12).from_address(pin[i])
            dst = (C.c_uint8 * 12).from_address(pout[i])
            vals = [(v + 1) % 256 for v in src]
            for j in range(12):
                dst[j] = vals[j]
        return 0

    def vsr_sttn_submit(self, h, pin, n, H, W, mask):
        self.tickets += 1
        self.calls.append(("submit", self.tickets))
        self._pending = getattr(self, "_pending", {})
        self._pending[self.tickets] = [bytes((C.c_uint8 * 12).from_address(pin[i])) for i in range(n)]
        return self.tickets

    def vsr_sttn_collect(self, h, ticket, pout):
        self.calls.append(("collect", ticket))
        if ticket in self.fail_collect:
            self.fail_collect.discard(ticket)
            return -5
        for i, src in enumerate(self._pending.pop(ticket)):     # the direct-mode result: in + 100 (must never survive a repeat)
            dst = (C.c_uint8 * 12).from_address(pout[i])
            for j in range(12):
                dst[j] = (src[j] + 100) % 256
        return 0


@pytest.fixture
def engine(monkeypatch):
    lib = FakeLib()
    monkeypatch.setattr(_capi, "lib", lambda: lib)
    eng = STTNInpaint.__new__(STTNInpaint)
    eng._h, eng._dev, eng.device = C.c_void_p(1), 0, "cuda:0"
    yield eng, lib
    eng._h = None          # nothing for __del__ to hand to the real library once the stand-in is gone


def _frames(n, base):
    return [np.full((2, 2, 3), base + i, np.uint8) for i in range(n)]


def test_synchronous_call_repeats_once_on_the_exact_path(engine):
    eng, lib = engine
    lib.fail_frames = {1}
    frames, mask = _frames(3, 10), np.zeros((2, 2), np.uint8)
    out = eng(frames, mask)
    assert lib.options == [("attn_direct", 0)] and lib.n_frames_calls == 2
    assert [int(o[0, 0, 0]) for o in out] == [11, 12, 13] and [int(f[0, 0, 0]) for f in frames] == [10, 11, 12]
    eng(frames, mask)                                   # later calls go straight through
    assert lib.n_frames_calls == 3 and len(lib.options) == 1


def test_pipeline_repeats_the_failed_chunk_and_the_one_in_flight(engine):
    eng, lib = engine
    mask = np.zeros((2, 2), np.uint8)
    a, b, c = _frames(2, 10), _frames(2, 50), _frames(2, 90)
    ta, tb = eng.submit(a, mask), eng.submit(b, mask)
    lib.fail_collect = {ta}
    eng.collect(ta, a)                                  # raises inside -> exact path, chunk a repeated from its untouched frames
    assert lib.options == [("attn_direct", 0)] and [int(f[0, 0, 0]) for f in a] == [11, 12]
    eng.collect(tb, b)                                  # b was in flight under the old mode: collected into scratch, repeated from ITS inputs
    assert [int(f[0, 0, 0]) for f in b] == [51, 52], "the direct-mode result (+100) or a double pass (+101) must not survive"
    tc = eng.submit(c, mask)
    eng.collect(tc, c)                                  # submitted after the switch: trusted (the stand-in paints +100 for every collect)
    assert [int(f[0, 0, 0]) for f in c] == [190, 191] and lib.n_frames_calls == 2
