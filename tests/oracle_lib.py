"""ctypes wrapper around the CPU oracle (oracle/fosphor_oracle.c).
TEST INFRASTRUCTURE ONLY - never imported by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libfosphor_oracle.so")


class Params(C.Structure):
    _fields_ = [("fft_len", C.c_int), ("n_bins", C.c_int), ("wf_rows", C.c_int),
                ("batch_mult", C.c_int), ("batch_max", C.c_int),
                ("histo_t0r", C.c_float), ("histo_t0d", C.c_float), ("live_alpha", C.c_float),
                ("maxhold_keep", C.c_float), ("maxhold_mix", C.c_float), ("fft_f32", C.c_int)]


def build_oracle(force=False):
    src = os.path.join(ORACLE_DIR, "fosphor_oracle.c")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return ORACLE_SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        vp, fp = C.c_void_p, C.POINTER(C.c_float)
        L.fosphor_oracle_default_params.argtypes = [C.POINTER(Params)]
        L.fosphor_oracle_create.argtypes = [C.POINTER(Params)]
        L.fosphor_oracle_create.restype = vp
        L.fosphor_oracle_destroy.argtypes = [vp]
        L.fosphor_oracle_load_fft_window.argtypes = [vp, vp]
        L.fosphor_oracle_default_window.argtypes = [C.c_int, vp]
        L.fosphor_oracle_power_range.argtypes = [C.c_int, C.c_int, C.c_int, fp, fp]
        L.fosphor_oracle_set_histogram_range.argtypes = [vp, C.c_float, C.c_float]
        L.fosphor_oracle_process.argtypes = [vp, vp, C.c_int]
        L.fosphor_oracle_process_hop.argtypes = [vp, vp, C.c_int, C.c_int]
        L.fosphor_oracle_process_pwr.argtypes = [vp, vp, C.c_int]
        L.fosphor_oracle_finish.argtypes = [vp]
        L.fosphor_oracle_get_waterfall_position.argtypes = [vp]
        for n in ("waterfall", "histogram", "spectrum", "last_fft", "last_hits"):
            getattr(L, "fosphor_oracle_" + n).argtypes = [vp]
            getattr(L, "fosphor_oracle_" + n).restype = vp
        L.fosphor_oracle_last_batch.argtypes = [vp]
        _lib = L
    return _lib


def default_window(n):
    w = np.empty(n, np.float32)
    lib().fosphor_oracle_default_window(n, w.ctypes.data)
    return w


def power_range(n, db_ref, db_per_div):
    s, o = C.c_float(), C.c_float()
    lib().fosphor_oracle_power_range(n, db_ref, db_per_div, C.byref(s), C.byref(o))
    return np.float32(s.value), np.float32(o.value)


class Oracle:
    def __init__(self, fft_len=1024, n_bins=128, wf_rows=1024, batch_mult=16, batch_max=1024,
                 t0r=16.0, t0d=1024.0, alpha=0.002, fft_f32=False, window=None,
                 db_ref=0, db_per_div=10):
        L = lib()
        p = Params()
        L.fosphor_oracle_default_params(C.byref(p))
        p.fft_len, p.n_bins, p.wf_rows = fft_len, n_bins, wf_rows
        p.batch_mult, p.batch_max = batch_mult, batch_max
        p.histo_t0r, p.histo_t0d, p.live_alpha = t0r, t0d, alpha
        p.fft_f32 = int(fft_f32)
        self.p = p
        self.n, self.k, self.w = fft_len, n_bins, wf_rows
        self.h = L.fosphor_oracle_create(C.byref(p))
        if not self.h:
            raise ValueError("bad oracle parameters")
        self.load_fft_window(default_window(fft_len) if window is None else window)
        self.set_power_range(db_ref, db_per_div)

    def load_fft_window(self, win):
        win = np.ascontiguousarray(win, np.float32)
        assert win.shape == (self.n,)
        lib().fosphor_oracle_load_fft_window(self.h, win.ctypes.data)

    def set_power_range(self, db_ref, db_per_div):
        s, o = power_range(self.n, db_ref, db_per_div)
        self.set_histogram_range(s, o)

    def set_histogram_range(self, scale, offset):
        lib().fosphor_oracle_set_histogram_range(self.h, float(scale), float(offset))

    def process(self, samples):
        x = np.ascontiguousarray(samples, np.complex64)
        return lib().fosphor_oracle_process(self.h, x.ctypes.data, x.size)

    def process_hop(self, raw, n_spectra, hop):
        x = np.ascontiguousarray(raw, np.complex64)
        assert (n_spectra - 1) * hop + self.n <= x.size or n_spectra == 0
        return lib().fosphor_oracle_process_hop(self.h, x.ctypes.data, n_spectra, hop)

    def process_pwr(self, rows):
        """one call on given log-power rows [B][N] (display stage only)"""
        r = np.ascontiguousarray(rows, np.float32)
        assert r.ndim == 2 and r.shape[1] == self.n
        return lib().fosphor_oracle_process_pwr(self.h, r.ctypes.data, r.shape[0])

    def finish(self):
        return lib().fosphor_oracle_finish(self.h)

    @property
    def waterfall_position(self):
        return lib().fosphor_oracle_get_waterfall_position(self.h)

    def _arr(self, name, shape, dtype=np.float32):
        ptr = getattr(lib(), "fosphor_oracle_" + name)(self.h)
        n = int(np.prod(shape))
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()

    @property
    def waterfall(self):
        return self._arr("waterfall", (self.w, self.n))

    @property
    def histogram(self):
        return self._arr("histogram", (self.k, self.n))

    @property
    def spectrum(self):
        return self._arr("spectrum", (2, self.n, 2))

    @property
    def last_fft(self):
        b = lib().fosphor_oracle_last_batch(self.h)
        return self._arr("last_fft", (b, self.n, 2)).view(np.complex64).reshape(b, self.n)

    @property
    def last_hits(self):
        return self._arr("last_hits", (self.k, self.n), np.uint32)

    def close(self):
        if self.h:
            lib().fosphor_oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
