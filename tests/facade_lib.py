"""ctypes driver for oracle/_ref/libfosphor_facade_b200.so: the reference's UNMODIFIED
lib/fosphor/fosphor.c linked against libfosphor_b200.so (oracle/ref_link/Makefile).
TEST INFRASTRUCTURE ONLY.  Drives the reference's public API (lib/fosphor/fosphor.h:26-37)
- fosphor_init / process / draw / set_fft_window / set_power_range / release - and exposes the
golden_cases.replay() interface, so the fixtures recorded from the real reference can be held
against "reference facade + CUDA drop-in" exactly as they are held against the drop-in alone."""
import ctypes as C
import os

import numpy as np

from gr_fosphor_b200.dropin import StructFosphor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FACADE_SO = os.path.join(ROOT, "oracle", "_ref", "libfosphor_facade_b200.so")


class _Channel(C.Structure):            # fosphor.h:44-49
    _fields_ = [("enabled", C.c_int), ("center", C.c_float), ("width", C.c_float)]


class Render(C.Structure):              # fosphor.h:61-91
    _fields_ = [("pos_x", C.c_int), ("pos_y", C.c_int), ("width", C.c_int), ("height", C.c_int),
                ("options", C.c_int), ("histo_wf_ratio", C.c_float), ("freq_n_div", C.c_int),
                ("freq_center", C.c_float), ("freq_span", C.c_float), ("wf_span", C.c_float),
                ("channels", _Channel * 8),
                ("_wf_pos", C.c_int), ("_x_div", C.c_float), ("_x", C.c_float * 2), ("_x_label", C.c_float),
                ("_y_histo_div", C.c_float), ("_y_histo", C.c_float * 2), ("_y_wf", C.c_float * 2),
                ("_y_label", C.c_float)]


class FosphorFacade:
    def __init__(self, path=FACADE_SO):
        L = self.lib = C.CDLL(path)
        P = C.POINTER(StructFosphor)
        L.fosphor_init.restype = P
        L.fosphor_release.argtypes = [P]
        L.fosphor_release.restype = None
        L.fosphor_process.argtypes = [P, C.c_void_p, C.c_int]
        L.fosphor_draw.argtypes = [P, C.POINTER(Render)]
        L.fosphor_draw.restype = None
        L.fosphor_set_fft_window.argtypes = [P, C.c_void_p]
        L.fosphor_set_fft_window.restype = None
        L.fosphor_set_power_range.argtypes = [P, C.c_int, C.c_int]
        L.fosphor_set_power_range.restype = None
        L.fosphor_render_defaults.argtypes = [C.POINTER(Render)]
        L.fosphor_render_defaults.restype = None
        L.fosphor_stub_refreshes.argtypes = [P]
        L.fosphor_stub_draws.argtypes = [P]
        self.s = L.fosphor_init()           # GL stub, then fosphor_cl_init of the drop-in, host images
        if not self.s:
            raise RuntimeError("fosphor_init() failed (no CUDA device?)")
        self.render = Render()
        L.fosphor_render_defaults(C.byref(self.render))

    def _img(self, ptr, shape):
        n = int(np.prod(shape))
        return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shape)

    @property
    def img_waterfall(self):
        return self._img(self.s.contents.img_waterfall, (1024, 1024))

    @property
    def img_histogram(self):
        return self._img(self.s.contents.img_histogram, (128, 1024))

    @property
    def buf_spectrum(self):
        return self._img(self.s.contents.buf_spectrum, (2, 1024, 2))

    def set_fft_window(self, win):
        w = np.ascontiguousarray(win, np.float32)
        self.lib.fosphor_set_fft_window(self.s, w.ctypes.data)      # copies into self->fft_win (fosphor.c:123-128)

    def set_power_range(self, db_ref, db_per_div):
        self.lib.fosphor_set_power_range(self.s, int(db_ref), int(db_per_div))

    def process(self, samples):
        x = np.ascontiguousarray(samples, np.complex64)
        return self.lib.fosphor_process(self.s, x.ctypes.data, x.size)

    def finish(self):
        """fosphor_draw(): fosphor_cl_finish() > 0 -> fosphor_gl_refresh(); 1 iff the renderer was refreshed"""
        before = self.lib.fosphor_stub_refreshes(self.s)
        self.lib.fosphor_draw(self.s, C.byref(self.render))
        return self.lib.fosphor_stub_refreshes(self.s) - before

    @property
    def waterfall_position(self):
        return int(self.render._wf_pos)                              # fosphor.c:103

    def release(self):
        if self.s:
            self.lib.fosphor_release(self.s)
            self.s = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass
