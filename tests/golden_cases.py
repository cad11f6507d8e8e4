"""Call sequences used to pin the oracle against the REAL reference.

Each case is a list of steps executed through the ``fosphor_cl_*`` boundary at
the reference's fixed problem size (N=1024, 128 bins, 1024 waterfall rows):
    ("window", array)            fosphor_set_fft_window
    ("range", db_ref, db_per_div) fosphor_set_power_range
    ("process", samples)         fosphor_process   (rc recorded)
    ("finish",)                  fosphor_cl_finish (rc + snapshot recorded)
tests/golden/make_golden.py replays them on the reference's own OpenCL path
(oracle/_ref/libfosphor_ref.so, GPU box) and stores the snapshots;
tests/test_oracle_golden.py replays them on the CPU oracle and
tests/test_dropin_parity.py on the CUDA drop-in.
"""
import numpy as np

import signals

N = 1024


def _rect():
    return np.ones(N, np.float32)


def _blackman_harris():
    n = np.arange(N, dtype=np.float64)
    a = (0.35875, 0.48829, 0.14128, 0.01168)
    w = a[0] - a[1] * np.cos(2 * np.pi * n / N) + a[2] * np.cos(4 * np.pi * n / N) - a[3] * np.cos(6 * np.pi * n / N)
    return w.astype(np.float32)


def cases():
    c = {}
    # BASELINE.json configs[0]: one 64k burst, default window + range
    c["cfg1"] = [("process", signals.cfg1_burst()), ("finish",)]

    # state carried over calls of different batch sizes, waterfall ring wrap
    stream = signals.noise_tones(N * (16 + 64 + 256 + 1024 + 32 + 1024), seed=7)
    steps, pos = [("finish",)], 0  # finish before any data: BOOTING -> clears
    for b in (16, 64, 256, 1024, 32, 1024):
        steps.append(("process", stream[pos:pos + b * N]))
        steps.append(("finish",))
        pos += b * N
    c["sequence"] = steps

    # several process calls per finish, like base_sink_c_impl::render()
    stream = signals.noise_tones(N * 8 * 128, seed=8, sigma=0.02)
    steps = []
    for i in range(8):
        steps.append(("process", stream[i * 128 * N:(i + 1) * 128 * N]))
    steps += [("finish",), ("finish",)]
    c["frame8"] = steps

    # analytic KATs with a rectangular window
    c["kat_tones"] = [
        ("window", _rect()),
        ("process", signals.tone(N, 16, 100, 1.0)), ("finish",),
        ("process", signals.tone(N, 16, 300, 0.1)), ("finish",),
        ("process", signals.tone(N, 16, 700, 0.01)), ("finish",),
        ("process", signals.impulse(N, 64)), ("finish",),
    ]

    # non-default window and power range
    c["bh_range"] = [
        ("window", _blackman_harris()),
        ("range", -10, 5),
        ("process", signals.noise_tones(N * 128, seed=9, sigma=0.003)), ("finish",),
        ("range", 10, 12),
        ("process", signals.noise_tones(N * 128, seed=10, sigma=0.05)), ("finish",),
    ]

    # zeros -> log10(0) = -inf through waterfall / live / max and the
    # !isfinite fallbacks on the following call (display.cl:206-207,290-291)
    c["zeros_then_data"] = [
        ("process", np.zeros(N * 16, np.complex64)), ("finish",),
        ("process", signals.noise_tones(N * 32, seed=11)), ("finish",),
    ]

    # argument validation (cl.c:881-886)
    c["einval"] = [
        ("process", np.zeros(N * 15, np.complex64)),
        ("process", np.zeros(N * 1040, np.complex64)),
        ("process", signals.noise_tones(N * 16, seed=12)), ("finish",),
    ]
    return c


def replay(eng, steps):
    """Run steps on any object with the FosphorCL / Oracle-adapter interface.
    Returns list of records: dicts with rc / wf_pos / arrays per finish."""
    out = []
    for st in steps:
        if st[0] == "window":
            eng.set_fft_window(st[1])
        elif st[0] == "range":
            eng.set_power_range(st[1], st[2])
        elif st[0] == "process":
            out.append({"op": "process", "rc": int(eng.process(st[1]))})
        elif st[0] == "finish":
            rc = int(eng.finish())
            out.append({"op": "finish", "rc": rc, "wf_pos": int(eng.waterfall_position),
                        "waterfall": np.array(eng.img_waterfall, copy=True),
                        "histogram": np.array(eng.img_histogram, copy=True),
                        "spectrum": np.array(eng.buf_spectrum, copy=True)})
    return out


class OracleAdapter:
    """Gives tests/oracle_lib.Oracle the FosphorCL interface (host result
    arrays refreshed only when finish() returns 1, like cl.c:1012-1048)."""

    def __init__(self):
        import oracle_lib
        self.o = oracle_lib.Oracle()
        self.img_waterfall = np.zeros((1024, N), np.float32)
        self.img_histogram = np.zeros((128, N), np.float32)
        self.buf_spectrum = np.zeros((2, N, 2), np.float32)

    def set_fft_window(self, w):
        self.o.load_fft_window(w)

    def set_power_range(self, a, b):
        self.o.set_power_range(a, b)

    def process(self, x):
        return self.o.process(x)

    def finish(self):
        rc = self.o.finish()
        if rc > 0:
            self.img_waterfall[:] = self.o.waterfall
            self.img_histogram[:] = self.o.histogram
            self.buf_spectrum[:] = self.o.spectrum
        return rc

    @property
    def waterfall_position(self):
        return self.o.waterfall_position
