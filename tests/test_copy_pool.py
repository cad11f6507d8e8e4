"""The staging copy pool (gr-fosphor_b200/host/copy_pool.{h,cc}) moves pageable caller memory into
page-locked slots - and results out of the bounce buffer - with a few persistent threads and
in-order pieces.  Host-only code: exercised here without a GPU through the library's self-test
hook (sizes around the piece / item boundaries, 1..8 workers, sleeping and polling workers)."""
import ctypes as C

import pytest


def _lib():
    from gr_fosphor_b200 import build
    L = C.CDLL(build.build())
    L.fosphor_host_copy_selftest.argtypes = [C.c_int, C.c_ulonglong, C.c_int, C.c_int]
    return L


@pytest.mark.parametrize("threads", [1, 2, 4, 8])
@pytest.mark.parametrize("nbytes,pieces", [(8 << 20, 2), (8 << 20, 8), ((8 << 20) + 4097, 5), (1234567, 3),
                                           (70000, 4), (4096, 1), (1, 1), ((1 << 20) - 1, 16)])
def test_pool_copies_exactly(threads, nbytes, pieces):
    assert _lib().fosphor_host_copy_selftest(threads, nbytes, pieces, 12) == 0


def test_default_thread_count_is_sane():
    assert _lib().fosphor_host_copy_selftest(0, 3 << 20, 3, 4) == 0      # 0 = automatic
