"""Host-side mirror of the reference's libfosphor facade over the ``fosphor_cl_*``
boundary (lib/fosphor/cl.h:22-32), driven through ctypes.

The same class drives BOTH shared objects that export those seven symbols:

* ``gr-fosphor_b200/libfosphor_b200.so``  - this repo's sm_100a CUDA drop-in
* ``oracle/_ref/libfosphor_ref.so``       - the reference's own, unmodified
  ``cl.c`` + OpenCL programs (test infrastructure, see oracle/ref_build/)

so a parity test is literally "run the same calls against both libraries".

What is mirrored (reference file:line):
  struct fosphor layout            lib/fosphor/private.h:30-55
  fosphor_init / release           lib/fosphor/fosphor.c:29-88
  fosphor_process                  lib/fosphor/fosphor.c:92-96
  fosphor_draw (compute half)      lib/fosphor/fosphor.c:98-105
  default window                   lib/fosphor/fosphor.c:108-121
  set_fft_window                   lib/fosphor/fosphor.c:123-128
  set_power_range                  lib/fosphor/fosphor.c:131-152
No GL: the renderer is out of scope, results stay in the host arrays the
reference's ``fosphor_gl_refresh`` would upload (gl.c:342-349).
"""
import ctypes as C

import numpy as np

FOSPHOR_FFT_LEN = 1024          # private.h:21-22
FOSPHOR_FFT_MULT_BATCH = 16     # private.h:24
FOSPHOR_FFT_MAX_BATCH = 1024    # private.h:25
FOSPHOR_N_BINS = 128            # display.cl:96, fosphor.c:53
FOSPHOR_WF_ROWS = 1024          # fosphor.c:52
FLG_FOSPHOR_USE_CLGL_SHARING = 1 << 0


class _Power(C.Structure):
    _fields_ = [("db_ref", C.c_int), ("db_per_div", C.c_int),
                ("scale", C.c_float), ("offset", C.c_float)]


class _Frequency(C.Structure):
    _fields_ = [("center", C.c_double), ("span", C.c_double)]


class StructFosphor(C.Structure):
    """``struct fosphor`` (private.h:30-55)."""
    _fields_ = [
        ("cl", C.c_void_p),
        ("gl", C.c_void_p),
        ("flags", C.c_int),
        ("fft_win", C.c_float * FOSPHOR_FFT_LEN),
        ("img_waterfall", C.POINTER(C.c_float)),
        ("img_histogram", C.POINTER(C.c_float)),
        ("buf_spectrum", C.POINTER(C.c_float)),
        ("power", _Power),
        ("frequency", _Frequency),
    ]


def default_window(n=FOSPHOR_FFT_LEN):
    """fosphor.c:113-118 in f32: periodic Hamming x 1.855, pi = 3.141592f."""
    f = np.float32
    i = np.arange(n, dtype=np.float32)
    arg = (f(2.0) * f(3.141592) * i) / f(n)
    return ((f(0.54) - f(0.46) * np.cos(arg, dtype=np.float32)) * f(1.855)).astype(np.float32)


def power_range(db_ref, db_per_div, n=FOSPHOR_FFT_LEN):
    """fosphor.c:131-152 -> (scale, offset) as f32."""
    f = np.float32
    db0 = db_ref - 10 * db_per_div
    db1 = db_ref
    k = f(np.log10(np.float64(n)))   # == C log10f((float)N): correctly rounded for these N
    offset = f(-(k + f(db0) / f(20.0)))
    scale = f(f(20.0) / f(db1 - db0))
    return scale, offset


class FosphorCL:
    """libfosphor facade over a library exporting ``fosphor_cl_*``."""

    def __init__(self, libpath):
        self.lib = C.CDLL(libpath)
        L = self.lib
        P = C.POINTER(StructFosphor)
        L.fosphor_cl_init.argtypes = [P]
        L.fosphor_cl_init.restype = C.c_int
        L.fosphor_cl_release.argtypes = [P]
        L.fosphor_cl_release.restype = None
        L.fosphor_cl_process.argtypes = [P, C.c_void_p, C.c_int]
        L.fosphor_cl_process.restype = C.c_int
        L.fosphor_cl_finish.argtypes = [P]
        L.fosphor_cl_finish.restype = C.c_int
        L.fosphor_cl_load_fft_window.argtypes = [P, C.POINTER(C.c_float)]
        L.fosphor_cl_load_fft_window.restype = None
        L.fosphor_cl_get_waterfall_position.argtypes = [P]
        L.fosphor_cl_get_waterfall_position.restype = C.c_int
        L.fosphor_cl_set_histogram_range.argtypes = [P, C.c_float, C.c_float]
        L.fosphor_cl_set_histogram_range.restype = None

        # fosphor_init(), fosphor.c:29-80
        self._inflight = []
        self.s = StructFosphor()
        C.memset(C.byref(self.s), 0, C.sizeof(self.s))
        rv = L.fosphor_cl_init(C.byref(self.s))
        if rv:
            raise RuntimeError("fosphor_cl_init failed: %d" % rv)
        if self.s.flags & FLG_FOSPHOR_USE_CLGL_SHARING:
            raise RuntimeError("CL/GL sharing requested but no GL exists in this harness")
        n = FOSPHOR_FFT_LEN
        self.img_waterfall = np.zeros((FOSPHOR_WF_ROWS, n), np.float32)
        self.img_histogram = np.zeros((FOSPHOR_N_BINS, n), np.float32)
        self.buf_spectrum = np.zeros((2, n, 2), np.float32)
        fp = C.POINTER(C.c_float)
        self.s.img_waterfall = self.img_waterfall.ctypes.data_as(fp)
        self.s.img_histogram = self.img_histogram.ctypes.data_as(fp)
        self.s.buf_spectrum = self.buf_spectrum.ctypes.data_as(fp)
        self.set_fft_window(default_window())
        self.set_power_range(0, 10)

    # fosphor.c:123-128
    def set_fft_window(self, win):
        win = np.ascontiguousarray(win, np.float32)
        assert win.shape == (FOSPHOR_FFT_LEN,)
        C.memmove(self.s.fft_win, win.ctypes.data, 4 * FOSPHOR_FFT_LEN)
        self.lib.fosphor_cl_load_fft_window(C.byref(self.s), self.s.fft_win)

    # fosphor.c:131-152
    def set_power_range(self, db_ref, db_per_div):
        scale, offset = power_range(db_ref, db_per_div)
        self.s.power.db_ref = db_ref
        self.s.power.db_per_div = db_per_div
        self.s.power.scale = float(scale)
        self.s.power.offset = float(offset)
        self.lib.fosphor_cl_set_histogram_range(C.byref(self.s), float(scale), float(offset))

    # fosphor.c:92-96
    def process(self, samples):
        samples = np.ascontiguousarray(samples, np.complex64)
        # the reference's H2D is non-blocking (cl.c:903-910): keep the buffer
        # alive until the next finish()
        self._inflight.append(samples)
        return self.process_raw(samples.ctypes.data, samples.size)

    def process_raw(self, ptr, length):
        return self.lib.fosphor_cl_process(C.byref(self.s), C.c_void_p(ptr), int(length))

    # compute half of fosphor_draw(), fosphor.c:98-105
    def finish(self):
        rv = self.lib.fosphor_cl_finish(C.byref(self.s))
        self._inflight = []
        return rv

    @property
    def waterfall_position(self):
        return self.lib.fosphor_cl_get_waterfall_position(C.byref(self.s))

    # fosphor.c:83-88
    def release(self):
        if self.s is not None:
            self.lib.fosphor_cl_release(C.byref(self.s))
            self.s = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass
