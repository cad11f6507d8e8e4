"""Host-side logic of the page-locked FIFO (gr-fosphor_b200/host/pinned_fifo.*),
the sink-side ring mirroring the reference's lib/fifo.{h,cc}.  Runs without a
GPU (the ring then falls back to ordinary memory)."""
import ctypes as C
import threading

import numpy as np


def _lib():
    from gr_fosphor_b200 import build
    L = C.CDLL(build.build())
    L.fosphor_fifo_create.restype = C.c_void_p
    L.fosphor_fifo_create.argtypes = [C.c_int]
    for n in ("destroy", "write_commit", "read_discard"):
        getattr(L, "fosphor_fifo_" + n).restype = None
    L.fosphor_fifo_destroy.argtypes = [C.c_void_p]
    for n in ("is_pinned", "free", "used", "write_max_size", "read_max_size"):
        f = getattr(L, "fosphor_fifo_" + n)
        f.argtypes, f.restype = [C.c_void_p], C.c_int
    for n in ("write_prepare", "read_peek"):
        f = getattr(L, "fosphor_fifo_" + n)
        f.argtypes, f.restype = [C.c_void_p, C.c_int, C.c_int], C.c_void_p
    L.fosphor_fifo_write_commit.argtypes = [C.c_void_p, C.c_int]
    L.fosphor_fifo_read_discard.argtypes = [C.c_void_p, C.c_int]
    return L


def test_bad_length_rejected():
    L = _lib()
    assert not L.fosphor_fifo_create(1000)       # not a power of two
    assert not L.fosphor_fifo_create(1)


def test_accounting_like_reference_ring():
    """lib/fifo.cc:28-45: free = len-1-used, contiguous sizes run to the ring end."""
    L = _lib()
    f = L.fosphor_fifo_create(1024)
    assert f
    assert L.fosphor_fifo_used(f) == 0 and L.fosphor_fifo_free(f) == 1023
    assert L.fosphor_fifo_write_max_size(f) == 1024 and L.fosphor_fifo_read_max_size(f) == 1024
    assert L.fosphor_fifo_read_peek(f, 1, 0) is None           # empty, no wait
    p = L.fosphor_fifo_write_prepare(f, 600, 0)
    assert p
    L.fosphor_fifo_write_commit(f, 600)
    assert L.fosphor_fifo_used(f) == 600 and L.fosphor_fifo_free(f) == 423
    assert L.fosphor_fifo_write_max_size(f) == 424
    assert L.fosphor_fifo_write_prepare(f, 424, 0) is None     # would need the reserved slot
    q = L.fosphor_fifo_read_peek(f, 600, 0)
    assert q == p
    L.fosphor_fifo_read_discard(f, 600)
    assert L.fosphor_fifo_used(f) == 0 and L.fosphor_fifo_read_max_size(f) == 424
    L.fosphor_fifo_destroy(f)


def test_producer_consumer_threads_preserve_the_stream():
    L = _lib()
    n, chunk, total = 4096, 256, 256 * 200
    f = L.fosphor_fifo_create(n)
    src = (np.arange(total) + 1j * np.arange(total)[::-1]).astype(np.complex64)
    out = np.empty(total, np.complex64)

    def producer():
        pos = 0
        while pos < total:
            m = min(chunk, L.fosphor_fifo_write_max_size(f), total - pos)
            dst = L.fosphor_fifo_write_prepare(f, m, 1)       # blocks while full
            C.memmove(dst, src[pos:pos + m].ctypes.data, 8 * m)
            L.fosphor_fifo_write_commit(f, m)
            pos += m

    t = threading.Thread(target=producer)
    t.start()
    pos = 0
    while pos < total:
        m = min(chunk, L.fosphor_fifo_read_max_size(f), total - pos)
        p = L.fosphor_fifo_read_peek(f, m, 1)                  # blocks while empty
        C.memmove(out[pos:pos + m].ctypes.data, p, 8 * m)
        L.fosphor_fifo_read_discard(f, m)
        pos += m
    t.join()
    assert np.array_equal(out, src)
    L.fosphor_fifo_destroy(f)
