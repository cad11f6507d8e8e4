#!/usr/bin/env python3
"""Host-fed drop-in throughput (sink frames: 8 x fosphor_cl_process(1024 spectra) + finish, pageable
numpy input) under the staging knobs of the engine, to pick its defaults.  One line per variant:
Msamples/s with the per-frame finish, and with one finish at the very end (the steady call rate)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def run(env, frames=150):
    import signals
    from gr_fosphor_b200 import build
    from gr_fosphor_b200.dropin import FosphorCL
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        eng = FosphorCL(build.LIB)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    n, b, calls = 1024, 1024, 8
    x = run.x
    call_len = b * n
    res = []
    for with_finish in (True, False):
        def frame():
            for c in range(calls):
                assert eng.process_raw(x.ctypes.data + 8 * c * call_len, call_len) == 0
            if with_finish:
                assert eng.finish() == 1
        for _ in range(5):
            frame()
        eng.finish()
        t0 = time.perf_counter()
        for _ in range(frames):
            frame()
        eng.finish()
        el = time.perf_counter() - t0
        res.append(frames * calls * call_len / el / 1e6)
    eng.release()
    return res


def main():
    import signals
    raw = signals.noise_tones((8 * 1024 - 1) * 256 + 1024, seed=2)
    run.x = signals.overlap_windows(raw, 1024, 4, 8 * 1024)
    variants = [{}, {"COPY_NT": "0"}, {"COPY_THREADS": "4"}, {"COPY_THREADS": "6"}, {"COPY_THREADS": "12"},
                {"STAGE_PIECE_KB": "1024"}, {"STAGE_PIECE_KB": "4096"}, {"STAGE_PIECE_KB": "8192"},
                {"STAGE_SLOTS": "2"}, {"STAGE_SLOTS": "8"}, {"HOSTREG": "1"}, {}]
    for v in variants:
        env = {"FOSPHOR_B200_" + k: val for k, val in v.items()}
        a, b = run(env)
        print("%-28s %8.0f Msamples/s with finish per frame   %8.0f Msamples/s calls only" % (v or "default", a, b), flush=True)


if __name__ == "__main__" and "phases" not in sys.argv:
    main()


def phases(frames=150):
    """Where a frame's time goes: the same frames through the parameterised engine (same geometry
    as the drop-in), with a sync between the calls and the read-back."""
    from gr_fosphor_b200.engine import Fosphor
    n, b, calls = 1024, 1024, 8
    x = run.x
    call_len = b * n
    eng = Fosphor(fft_len=n, n_bins=128, wf_rows=1024, scratch_rows=-1)
    t_calls = t_sync = t_fin = 0.0
    for f in range(frames + 5):
        t0 = time.perf_counter()
        for c in range(calls):
            eng.process_host_ptr(x.ctypes.data + 8 * c * call_len, call_len)
        t1 = time.perf_counter()
        eng.sync()
        t2 = time.perf_counter()
        rc, host, r0, nr = eng.finish_new_rows()
        t3 = time.perf_counter()
        if f >= 5:
            t_calls += t1 - t0
            t_sync += t2 - t1
            t_fin += t3 - t2
    print("phases per frame: 8 process calls %.0f us, wait for the GPU %.0f us, read-back of %d rows + histogram + "
          "spectrum into pageable arrays %.0f us" % (t_calls / frames * 1e6, t_sync / frames * 1e6, nr, t_fin / frames * 1e6))
    eng.close()


if __name__ == "__main__" and "phases" in sys.argv:
    import signals as _s
    _raw = _s.noise_tones((8 * 1024 - 1) * 256 + 1024, seed=2)
    run.x = _s.overlap_windows(_raw, 1024, 4, 8 * 1024)
    for _ in range(3):
        phases()
