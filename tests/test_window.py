"""Window generator (gr-fosphor_b200/host/window.cc) against
scipy's symmetric windows where the definitions coincide with GNU Radio's."""
import ctypes as C

import numpy as np
import pytest
from scipy import signal as sps


def _build(t, n, beta=6.76):
    from gr_fosphor_b200 import build
    L = C.CDLL(build.build())
    L.fosphor_window_build.argtypes = [C.c_int, C.c_int, C.c_double, C.c_void_p]
    w = np.zeros(n, np.float32)
    rc = L.fosphor_window_build(t, n, beta, w.ctypes.data)
    return rc, w


@pytest.mark.parametrize("t,name", [(0, "hamming"), (1, "hann"), (2, "blackman"), (3, "boxcar"),
                                    (5, "blackmanharris"), (6, "bartlett")])
def test_cosine_windows_match_scipy(t, name):
    rc, w = _build(t, 1024)
    assert rc == 0
    assert np.abs(w - sps.get_window(name, 1024, fftbins=False)).max() < 2e-7


def test_kaiser_and_flattop():
    rc, w = _build(4, 1024, 6.76)           # the sink's beta, base_sink_c_impl.cc:253
    assert rc == 0 and np.abs(w - sps.get_window(("kaiser", 6.76), 1024, fftbins=False)).max() < 2e-7
    rc, w = _build(7, 1024)
    assert rc == 0 and abs(w.max() - 1.0) < 1e-3 and w.min() < 0   # flat-top dips negative
    assert np.allclose(w, w[::-1], atol=1e-7)


def test_bad_arguments():
    assert _build(99, 1024)[0] == -1
    assert _build(0, 1)[0] == -1
