"""CPU: host-side logic that needs no GPU - the roofline byte counts of bench.py against SURVEY.md section 8(d), the
packed-buffer layout of _native.PackedBuffers, and the link-arbitration locks."""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_algorithmic_bytes_match_survey_table():
    import bench
    # fwd / fwd+bwd bytes per solve for c1/c4, c2, c3, c5
    want = {(3, 1, 20): (6144, 17952), (4, 2, 50): (30208, 89024), (8, 4, 50): (108032, 320896),
            (32, 8, 100): (2383360, 7124480)}
    for (n, m, T), (fwd, tot) in want.items():
        assert bench.algorithmic_bytes(n, m, T) == (fwd, tot)


class _FakeCtx:
    """Just enough of _native.Context for DeviceArray / PackedBuffers bookkeeping."""

    def __init__(self):
        self.h = None
        self.next = 1 << 20

    def _alloc(self, nbytes):
        p = self.next
        self.next += (nbytes + 255) // 256 * 256
        return p

    def _release(self, ptr, nbytes):
        pass

    def pinned_empty(self, shape, dtype=np.float64):      # the real context hands out page-locked memory
        return np.empty(shape, dtype)


def test_packed_buffers_layout_is_aligned_and_disjoint():
    import _native
    specs = [("x", (20, 64, 3), np.float64), ("free", (20, 64, 1), np.uint8), ("n_qp", (20, 64), np.int32),
             ("costs", (64,), np.float64), ("empty", (0, 4), np.float64)]
    pb = _native.PackedBuffers(_FakeCtx(), specs)
    spans = []
    for name, shape, dt in specs:
        v = pb.views[name]
        off = v.ptr - pb.base.ptr
        assert off % _native.PackedBuffers.ALIGN == 0 and v.shape == tuple(shape) and v.dtype == np.dtype(dt)
        assert v.nbytes == int(np.prod(shape)) * np.dtype(dt).itemsize and not v._owned
        spans.append((off, off + v.nbytes))
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
    assert spans[-1][1] <= pb.total
    # acquire / release keep one instance per layout in the context
    ctx = _FakeCtx()
    a = _native.PackedBuffers.acquire(ctx, specs)
    a.release()
    b = _native.PackedBuffers.acquire(ctx, specs)
    assert b is a
    c = _native.PackedBuffers.acquire(ctx, specs)          # a second live user gets its own buffers
    assert c is not a


def test_link_lock_is_per_device_and_direction_and_serialises_bursts():
    import _native
    assert _native.link_lock(0, "h2d") is _native.link_lock(0, "h2d")
    assert _native.link_lock(0, "h2d") is not _native.link_lock(0, "d2h")
    assert _native.link_lock(0, "h2d") is not _native.link_lock(1, "h2d")
    inside, worst = [0], [0]

    def burst():
        with _native.link_lock(3, "h2d"):
            inside[0] += 1
            worst[0] = max(worst[0], inside[0])
            time.sleep(0.01)
            inside[0] -= 1
    ths = [threading.Thread(target=burst) for _ in range(4)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert worst[0] == 1
