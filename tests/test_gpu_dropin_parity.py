"""GPU parity at the DROP-IN boundary: libfosphor_b200.so driven through the
reference's seven fosphor_cl_* entry points, against (a) the committed golden
vectors from the real reference and (b), when an OpenCL device is reachable,
the reference library itself running the same calls side by side."""
import os

import numpy as np
import pytest

import golden_cases
import golden_check
import parity

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libfosphor_ref.so")


def _dropin():
    from gr_fosphor_b200 import build
    from gr_fosphor_b200.dropin import FosphorCL
    return FosphorCL(build.LIB)


@pytest.mark.parametrize("name", ["cfg1", "sequence", "frame8", "kat_tones", "bh_range",
                                  "zeros_then_data", "einval"])
def test_dropin_matches_reference_golden(name):
    eng = _dropin()
    golden_check.check_case(name, eng)
    eng.release()


@pytest.mark.parametrize("name", ["cfg1", "sequence", "frame8", "bh_range", "einval"])
def test_reference_fosphor_c_over_the_dropin_matches_golden(name):
    """Link-level drop-in proof: the reference's unmodified fosphor.c (fosphor_init / process / draw,
    fosphor.h:26-37) linked against libfosphor_b200.so - oracle/ref_link - replays the golden
    cases; the img_* buffers fosphor.c owns (fosphor.c:52-54) must match what the real reference
    produced, fosphor_gl_refresh() must be requested exactly when the reference's finish said so."""
    import facade_lib
    if not os.path.exists(facade_lib.FACADE_SO):
        pytest.skip("oracle/_ref/libfosphor_facade_b200.so not built")
    eng = facade_lib.FosphorFacade()
    golden_check.check_case(name, eng)
    eng.release()


def test_dropin_side_by_side_with_live_reference():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libfosphor_ref.so not built")
    from gr_fosphor_b200.dropin import FosphorCL
    try:
        ref = FosphorCL(REF_SO)
    except RuntimeError:
        pytest.skip("no OpenCL device reachable for the reference library")
    mine = _dropin()
    steps = golden_cases.cases()["sequence"]
    a = golden_cases.replay(ref, steps)
    b = golden_cases.replay(mine, steps)
    spectra = 0
    for st, ra, rb in zip([s for s in steps if s[0] in ("process", "finish")], a, b):
        assert ra["rc"] == rb["rc"]
        if st[0] == "process":
            spectra += st[1].size // 1024
            continue
        assert ra["wf_pos"] == rb["wf_pos"]
        import oracle_lib
        sc, of = oracle_lib.power_range(1024, 0, 10)
        parity.check_end_to_end(rb, ra, np.arange(min(1024, max(1, spectra))), max(1, spectra) * 1024,
                                np.float32(sc) * np.float32(128), of)
    ref.release()
    mine.release()


def test_dropin_release_is_null_safe_and_idempotent():
    """cl.c:846-868: release with self->cl == NULL is a no-op; called twice on the
    init-failure path (fosphor.c:71-73,86)."""
    eng = _dropin()
    eng.release()
    import ctypes as C
    from gr_fosphor_b200.dropin import StructFosphor
    s = StructFosphor()
    eng.lib.fosphor_cl_release(C.byref(s))
    eng.lib.fosphor_cl_release(C.byref(s))
    assert not s.cl


def test_pinned_fifo_feeds_the_dropin_without_staging():
    """samples written into the page-locked FIFO are handed to
    fosphor_cl_process() straight from the ring (DMA, no staging memcpy) and
    discarded right after the call returns, like base_sink_c_impl::render()
    (lib/base_sink_c_impl.cc:146-175); results match the oracle."""
    import ctypes as C
    import oracle_lib
    import signals
    from test_pinned_fifo import _lib
    L = _lib()
    f = L.fosphor_fifo_create(1 << 21)              # the sink's 2 Mi-sample ring (base_sink_c_impl.cc:58)
    assert f and L.fosphor_fifo_is_pinned(f) == 1
    eng = _dropin()
    orc = oracle_lib.Oracle()
    x = signals.noise_tones(1024 * (256 + 1024 + 64), seed=23)
    pos = 0
    for b in (256, 1024, 64):
        n = b * 1024
        dst = L.fosphor_fifo_write_prepare(f, n, 1)
        C.memmove(dst, x[pos:pos + n].ctypes.data, 8 * n)
        L.fosphor_fifo_write_commit(f, n)
        src = L.fosphor_fifo_read_peek(f, n, 0)
        assert src
        assert eng.process_raw(src, n) == 0
        L.fosphor_fifo_read_discard(f, n)           # legal at once: the call has consumed the buffer
        C.memset(src, 0xff, 8 * n)                  # and the producer may overwrite it
        assert orc.process(x[pos:pos + n]) == 0
        pos += n
    assert eng.finish() == 1 and orc.finish() == 1
    sc, of = oracle_lib.power_range(1024, 0, 10)
    host = {"waterfall": eng.img_waterfall, "histogram": eng.img_histogram, "spectrum": eng.buf_spectrum}
    parity.check_end_to_end(host, orc, np.arange(1024), (256 + 1024 + 64) * 1024, np.float32(sc) * np.float32(128), of)
    eng.release()
    L.fosphor_fifo_destroy(f)
