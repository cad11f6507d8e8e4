"""Tolerance helpers shared by the parity tests.

Stated tolerances (floating-point path, the reference itself is not
bit-reproducible across OpenCL runtimes: native_sin/cos/powr, fft.cl:66-67,
display.cl:150,210,242-245):
  waterfall   |d pwr| <= 1e-4 log10 units, or |d mag| <= 1e-5 * max(mag) per row
              (bins far below the row maximum are ill-conditioned in log)
  histogram   |d hv| <= 2e-3 except for cells touched by a rounding-boundary
              bin flip; at most FLIP_FRAC of the call's hits may flip by +-1 bin
  live / max  |d| <= 1e-4
"""
import numpy as np

PWR_TOL = 1e-4
MAG_REL_TOL = 1e-5
HIST_TOL = 2e-3
SPEC_TOL = 1e-4
FLIP_FRAC = 0.005


def _eq_nonfinite(a, b):
    return (~np.isfinite(a)) & (~np.isfinite(b)) & ((a == b) | (np.isnan(a) & np.isnan(b)))


def check_waterfall(got, ref, rows=None):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    if rows is not None:
        got, ref = got[rows], ref[rows]
    same = _eq_nonfinite(got, ref)
    assert np.all(np.isfinite(got) | same), "non-finite waterfall values differ"
    d = np.where(same, 0.0, np.abs(got - ref))
    mag_g, mag_r = 10.0 ** np.where(same, 0, got), 10.0 ** np.where(same, 0, ref)
    rowmax = np.maximum(mag_r.max(axis=-1, keepdims=True), 1e-300)
    ok = (d <= PWR_TOL) | (np.abs(mag_g - mag_r) <= MAG_REL_TOL * rowmax)
    assert ok.all(), "waterfall: %d cells out of tolerance, worst |dpwr|=%g" % ((~ok).sum(), d[~ok].max())
    return float(d.max())


def check_histogram(got, ref, hits_in_play):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    d = np.abs(got - ref)
    bad = int((d > HIST_TOL).sum())
    budget = int(2 * FLIP_FRAC * hits_in_play) + 2
    assert bad <= budget, "histogram: %d cells differ by > %g (flip budget %d), worst %g" % (
        bad, HIST_TOL, budget, d.max())
    # a flipped hit moves mass to the neighbouring bin only: column sums stay close
    cs = np.abs(got.sum(axis=0) - ref.sum(axis=0))
    assert cs.max() <= 0.05 + 1e-3 * ref.sum(axis=0).max(), "histogram column mass differs: %g" % cs.max()
    return bad, float(d.max())


def conditioned_columns(wf_rows_ref, floor_db=80.0):
    """Columns whose magnitude stays within floor_db of the row maximum in every
    given waterfall row.  Bins that hold only rounding noise (e.g. the off-peak
    bins of a pure tone under a rectangular window) are ill-conditioned in log
    units in ANY f32 implementation, the reference included, and are excluded
    from the live / max-hold comparison.  Display order index i = f ^ N/2."""
    wf = np.asarray(wf_rows_ref, np.float64)
    rowmax = np.nanmax(np.where(np.isfinite(wf), wf, -np.inf), axis=1, keepdims=True)
    ok = np.all((wf >= rowmax - floor_db / 20.0) | ~np.isfinite(wf), axis=0)
    n = wf.shape[1]
    return ok[np.arange(n) ^ (n // 2)]


def check_spectrum(got, ref, cols=None):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    if cols is not None:
        got, ref = got[:, cols], ref[:, cols]
        if got.size == 0:
            return 0.0
    same = _eq_nonfinite(got, ref)
    d = np.where(same, 0.0, np.abs(got - ref))
    d = np.nan_to_num(d, nan=np.inf)
    assert d.max() <= SPEC_TOL, "spectrum: worst |d|=%g at %s" % (d.max(), np.unravel_index(d.argmax(), d.shape))
    return float(d.max())
