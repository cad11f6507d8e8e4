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


def test_flattop_pinned_to_the_published_five_term_set():
    """gr-fft's flat-top: the HP / SRS 5-term cosine sum {1, 1.93, 1.29, 0.388, 0.028} / 4.63867,
    symmetric over n - 1 intervals - evaluated here in double, every tap."""
    for n in (1024, 512, 4096):
        rc, w = _build(7, n)
        assert rc == 0
        c = np.array([1.0, 1.93, 1.29, 0.388, 0.028]) / 4.63867
        i = np.arange(n, dtype=np.float64)
        ref = sum((-1.0) ** k * c[k] * np.cos(2 * np.pi * k * i / (n - 1)) for k in range(5))
        assert np.abs(w - ref).max() < 1.5e-7
        assert abs(w.max() - 4.636 / 4.63867) < 1e-4            # peak = sum of the coefficients (n even: no tap at the exact centre)
        assert abs(float(w[0]) - (c[0] - c[1] + c[2] - c[3] + c[4])) < 1e-7
    # not the ISO 18431-2 / Matlab / scipy coefficient set: close, but a different window
    rc, w = _build(7, 1024)
    d = np.abs(w - sps.get_window("flattop", 1024, fftbins=False)).max()
    assert 2e-4 < d < 3e-3


@pytest.mark.parametrize("t", range(8))
def test_convention_is_symmetric_not_periodic(t):
    """gr::fft::window::build() windows span ntaps - 1 intervals: w[i] == w[n-1-i] exactly; a
    periodic ("fftbins") window of the same type would have w[1] == w[n-1] instead."""
    rc, w = _build(t, 1024)
    assert rc == 0 and np.array_equal(w, w[::-1])
    if t != 3:
        assert w[1] != w[0] and abs(float(w[1]) - float(w[-1])) > 0


def test_bad_arguments():
    assert _build(99, 1024)[0] == -1
    assert _build(0, 1)[0] == -1
