"""ctypes mirror of the parameterised ``fosphor_cu_*`` C ABI
(include/fosphor_b200.h) - test / bench harness only; the product is
``libfosphor_b200.so``.  There is no fallback: if the library is missing or no
CUDA device is usable, construction raises."""
import ctypes as C
import os

import numpy as np

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))


class Params(C.Structure):
    """``struct fosphor_cu_params``"""
    _fields_ = [("fft_len", C.c_int), ("n_bins", C.c_int), ("wf_rows", C.c_int),
                ("batch_mult", C.c_int), ("batch_max", C.c_int),
                ("histo_t0r", C.c_float), ("histo_t0d", C.c_float), ("live_alpha", C.c_float),
                ("maxhold_keep", C.c_float), ("maxhold_mix", C.c_float), ("device", C.c_int),
                ("scratch_rows", C.c_int)]


_lib = None


def load_library(path=None):
    """Load libfosphor_b200.so (the in-tree build).  Raises if it is absent:
    the product path never runs without the CUDA library."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or _build.LIB
    if not os.path.exists(path):
        raise RuntimeError("%s not built - run `python __graft_entry__.py` / build()" % path)
    L = C.CDLL(path)
    vp, ll = C.c_void_p, C.c_longlong
    L.fosphor_cu_default_params.argtypes = [C.POINTER(Params)]
    L.fosphor_cu_default_params.restype = None
    L.fosphor_cu_create.argtypes = [C.POINTER(vp), C.POINTER(Params)]
    L.fosphor_cu_destroy.argtypes = [vp]
    L.fosphor_cu_destroy.restype = None
    L.fosphor_cu_set_stream.argtypes = [vp, vp]
    L.fosphor_cu_load_fft_window.argtypes = [vp, vp]
    L.fosphor_cu_set_histogram_range.argtypes = [vp, C.c_float, C.c_float]
    L.fosphor_cu_process_host.argtypes = [vp, vp, C.c_int]
    L.fosphor_cu_process_device.argtypes = [vp, vp, C.c_int, ll]
    L.fosphor_cu_process_device_multi.argtypes = [vp, vp, C.c_int, C.c_int, ll]
    L.fosphor_cu_process_host_raw.argtypes = [vp, vp, C.c_int, C.c_int, ll]
    L.fosphor_cu_finish.argtypes = [vp, vp, vp, vp]
    L.fosphor_cu_finish_new_rows.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.fosphor_cu_default_window.argtypes = [C.c_int, vp]
    L.fosphor_cu_default_window.restype = None
    L.fosphor_cu_power_range.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.fosphor_cu_power_range.restype = None
    L.fosphor_cu_export_maxhold_on.argtypes = [vp, vp, vp]
    L.fosphor_cu_host_feed_stats.argtypes = [vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong),
                                             C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.fosphor_cu_sync.argtypes = [vp]
    L.fosphor_cu_flush.argtypes = [vp]
    L.fosphor_cu_two_stream_chunks.argtypes = [vp]
    L.fosphor_cu_two_stream_chunks.restype = C.c_ulonglong
    L.fosphor_cu_get_waterfall_position.argtypes = [vp]
    for n in ("waterfall", "histogram", "spectrum"):
        f = getattr(L, "fosphor_cu_device_" + n)
        f.argtypes = [vp]
        f.restype = vp
    L.fosphor_cu_export_maxhold.argtypes = [vp, vp]
    L.fosphor_cu_debug_fft.argtypes = [vp, vp, C.c_int, ll, vp]
    L.fosphor_cu_profile.argtypes = [vp, C.c_int]
    L.fosphor_cu_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)]
    L.fosphor_cu_launch_count.argtypes = [vp]
    L.fosphor_cu_launch_count.restype = C.c_ulonglong
    L.fosphor_cu_last_error.argtypes = [vp]
    L.fosphor_cu_last_error.restype = C.c_char_p
    if path == _build.LIB:
        _lib = L
    return L


def default_window(n):
    """lib/fosphor/fosphor.c:113-118 generalised to N: the library's C helper."""
    w = np.empty(n, np.float32)
    load_library().fosphor_cu_default_window(n, w.ctypes.data)
    return w


def power_range(n, db_ref, db_per_div):
    """lib/fosphor/fosphor.c:131-152 -> (scale, offset) f32: the library's C helper."""
    s, o = C.c_float(), C.c_float()
    load_library().fosphor_cu_power_range(n, int(db_ref), int(db_per_div), C.byref(s), C.byref(o))
    return np.float32(s.value), np.float32(o.value)


class Fosphor:
    """One engine instance == one fosphor display (one channel)."""

    def __init__(self, fft_len=1024, n_bins=128, wf_rows=1024, batch_mult=16, batch_max=1024,
                 t0r=16.0, t0d=1024.0, alpha=0.002, device=-1, window=None,
                 db_ref=0, db_per_div=10, stream=None, scratch_rows=0):
        self.lib = load_library()
        p = Params()
        self.lib.fosphor_cu_default_params(C.byref(p))
        p.fft_len, p.n_bins, p.wf_rows = fft_len, n_bins, wf_rows
        p.batch_mult, p.batch_max = batch_mult, batch_max
        p.histo_t0r, p.histo_t0d, p.live_alpha = t0r, t0d, alpha
        p.device = device
        p.scratch_rows = scratch_rows
        self.p = p
        self.n, self.k, self.w = fft_len, n_bins, wf_rows
        self.h = C.c_void_p()
        rc = self.lib.fosphor_cu_create(C.byref(self.h), C.byref(p))
        if rc:
            self.h = None
            raise RuntimeError("fosphor_cu_create failed: %d" % rc)
        if stream is not None:
            self.set_stream(stream)
        self.load_fft_window(default_window(fft_len) if window is None else window)
        self.set_power_range(db_ref, db_per_div)

    def _chk(self, rc):
        if rc < 0 and rc != -22:
            raise RuntimeError("fosphor_b200 error %d: %s" % (rc, self.lib.fosphor_cu_last_error(self.h).decode()))
        return rc

    def set_stream(self, stream):
        return self._chk(self.lib.fosphor_cu_set_stream(self.h, C.c_void_p(int(stream) if stream else 0)))

    def load_fft_window(self, win):
        win = np.ascontiguousarray(win, np.float32)
        assert win.shape == (self.n,)
        return self._chk(self.lib.fosphor_cu_load_fft_window(self.h, win.ctypes.data))

    def set_power_range(self, db_ref, db_per_div):
        s, o = power_range(self.n, db_ref, db_per_div)
        return self.set_histogram_range(s, o)

    def set_histogram_range(self, scale, offset):
        return self._chk(self.lib.fosphor_cu_set_histogram_range(self.h, float(scale), float(offset)))

    def process(self, samples):
        """Host cf32 samples, pre-overlapped windows (reference contract)."""
        x = np.ascontiguousarray(samples, np.complex64)
        return self._chk(self.lib.fosphor_cu_process_host(self.h, x.ctypes.data, x.size))

    def process_host_ptr(self, ptr, length):
        return self._chk(self.lib.fosphor_cu_process_host(self.h, C.c_void_p(ptr), int(length)))

    def process_host_raw(self, raw, n_calls, batch, hop):
        x = np.ascontiguousarray(raw, np.complex64)
        assert n_calls * batch == 0 or (n_calls * batch - 1) * hop + self.n <= x.size
        return self._chk(self.lib.fosphor_cu_process_host_raw(self.h, x.ctypes.data, n_calls, batch, hop))

    def process_host_raw_ptr(self, ptr, n_calls, batch, hop):
        return self._chk(self.lib.fosphor_cu_process_host_raw(self.h, C.c_void_p(ptr), n_calls, batch, hop))

    def process_device(self, dev_ptr, n_spectra, hop=None):
        return self._chk(self.lib.fosphor_cu_process_device(
            self.h, C.c_void_p(dev_ptr), n_spectra, self.n if hop is None else hop))

    def process_device_multi(self, dev_ptr, n_calls, batch, hop=None):
        return self._chk(self.lib.fosphor_cu_process_device_multi(
            self.h, C.c_void_p(dev_ptr), n_calls, batch, self.n if hop is None else hop))

    def finish(self, want=("waterfall", "histogram", "spectrum")):
        """Returns (rc, dict of host arrays); arrays only refreshed when rc == 1."""
        if not hasattr(self, "_host"):
            self._host = {"waterfall": np.zeros((self.w, self.n), np.float32),
                          "histogram": np.zeros((self.k, self.n), np.float32),
                          "spectrum": np.zeros((2, self.n, 2), np.float32)}
        ptrs = [self._host[k].ctypes.data if k in want else None
                for k in ("waterfall", "histogram", "spectrum")]
        rc = self._chk(self.lib.fosphor_cu_finish(self.h, *ptrs))
        return rc, self._host

    def finish_new_rows(self):
        """finish() that refreshes only the waterfall rows written since the previous one in the
        persistent host image; returns (rc, host arrays, first_row, n_rows)."""
        if not hasattr(self, "_host"):
            self._host = {"waterfall": np.zeros((self.w, self.n), np.float32),
                          "histogram": np.zeros((self.k, self.n), np.float32),
                          "spectrum": np.zeros((2, self.n, 2), np.float32)}
        r0, nr = C.c_int(), C.c_int()
        rc = self._chk(self.lib.fosphor_cu_finish_new_rows(
            self.h, self._host["waterfall"].ctypes.data, self._host["histogram"].ctypes.data,
            self._host["spectrum"].ctypes.data, C.byref(r0), C.byref(nr)))
        return rc, self._host, r0.value, nr.value

    def finish_into(self, wf_ptr, hist_ptr, spec_ptr):
        """finish() into caller-owned host memory (e.g. page-locked buffers); 0 pointers skip an array."""
        return self._chk(self.lib.fosphor_cu_finish(self.h, C.c_void_p(wf_ptr or None),
                                                    C.c_void_p(hist_ptr or None), C.c_void_p(spec_ptr or None)))

    def sync(self):
        return self._chk(self.lib.fosphor_cu_sync(self.h))

    @property
    def two_stream_chunks(self):
        return int(self.lib.fosphor_cu_two_stream_chunks(self.h))

    def flush(self):
        """order the engine's stream after all process work enqueued so far (no host wait)"""
        return self._chk(self.lib.fosphor_cu_flush(self.h))

    @property
    def waterfall_position(self):
        return self.lib.fosphor_cu_get_waterfall_position(self.h)

    @property
    def launch_count(self):
        return int(self.lib.fosphor_cu_launch_count(self.h))

    def profile(self, enable=True):
        return self._chk(self.lib.fosphor_cu_profile(self.h, int(enable)))

    def profile_read(self):
        ms = (C.c_double * 3)()
        nl = (C.c_ulonglong * 3)()
        self._chk(self.lib.fosphor_cu_profile_read(self.h, ms, nl))
        return {"fft_ms": ms[0], "fft_launches": nl[0], "count_ms": ms[1], "count_launches": nl[1],
                "update_ms": ms[2], "update_launches": nl[2]}

    def device_ptrs(self):
        return {n: getattr(self.lib, "fosphor_cu_device_" + n)(self.h)
                for n in ("waterfall", "histogram", "spectrum")}

    def export_maxhold(self, out_dev_ptr):
        return self._chk(self.lib.fosphor_cu_export_maxhold(self.h, C.c_void_p(out_dev_ptr)))

    def export_maxhold_on(self, out_dev_ptr, side_stream):
        """export on a caller stream (e.g. the NCCL stream) without joining the engine's streams"""
        return self._chk(self.lib.fosphor_cu_export_maxhold_on(self.h, C.c_void_p(out_dev_ptr),
                                                                C.c_void_p(int(side_stream))))

    def host_feed_stats(self):
        a, b, c, d = C.c_ulonglong(), C.c_ulonglong(), C.c_int(), C.c_int()
        self._chk(self.lib.fosphor_cu_host_feed_stats(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"staged_calls": a.value, "direct_calls": b.value, "copy_threads": c.value, "ring_rows": d.value}

    def debug_fft(self, in_dev_ptr, n_spectra, hop, out_dev_ptr):
        return self._chk(self.lib.fosphor_cu_debug_fft(
            self.h, C.c_void_p(in_dev_ptr), n_spectra, hop, C.c_void_p(out_dev_ptr)))

    def close(self):
        if getattr(self, "h", None):
            self.lib.fosphor_cu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
