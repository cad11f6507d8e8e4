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


FFT_REL = 1e-6      # relative (to the row maximum) error budget of an f32 FFT, averaged over a call


def column_depth(wf_rows_ref):
    """Per display-order column: log10 distance between the largest value seen in
    the given waterfall rows and the smallest value seen in that column.  A bin
    `depth` below the strongest line carries a relative rounding error of about
    FFT_REL * 10**depth in ANY f32 FFT (the reference's included), i.e.
    FFT_REL * 10**depth / ln(10) in log10 units."""
    wf = np.asarray(wf_rows_ref, np.float64)
    fin = np.isfinite(wf)
    top = np.max(np.where(fin, wf, -np.inf))
    colmin = np.min(np.where(fin, wf, np.inf), axis=0)
    colmin = np.where(np.isfinite(colmin), colmin, top)      # all -inf columns: compared exactly
    depth = np.maximum(top - colmin, 0.0)
    n = wf.shape[1]
    return depth[np.arange(n) ^ (n // 2)]


def conditioned_columns(wf_rows_ref, floor_db=80.0):
    """Columns that stay within floor_db of the strongest line.  Bins that hold
    only rounding noise (e.g. the off-peak bins of a pure tone under a
    rectangular window) are ill-conditioned in log units and are excluded from
    the live / max-hold comparison."""
    return column_depth(wf_rows_ref) <= floor_db / 20.0


def check_spectrum(got, ref, cols=None, wf_ref=None):
    """live / max-hold: |d| <= SPEC_TOL + FFT_REL * 10**depth / ln(10) per column
    (depth from wf_ref, see column_depth); columns deeper than 80 dB skipped."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    tol = np.full(got.shape[1], SPEC_TOL)
    if wf_ref is not None:
        depth = column_depth(wf_ref)
        tol = SPEC_TOL + FFT_REL * 10.0 ** np.minimum(depth, 4.0) / np.log(10.0)
        cols = (depth <= 4.0) if cols is None else (cols & (depth <= 4.0))
    if cols is not None:
        got, ref, tol = got[:, cols], ref[:, cols], tol[cols]
        if got.size == 0:
            return 0.0
    same = _eq_nonfinite(got, ref)
    d = np.where(same, 0.0, np.abs(got - ref))
    d = np.nan_to_num(d, nan=np.inf)
    excess = d[:, :, 1] - tol[None, :]
    assert np.all(d[:, :, 0] <= 1e-6), "x coordinates differ"
    assert excess.max() <= 0, "spectrum: worst |d|=%g (tol %g) at %s" % (
        d[:, :, 1].max(), tol.min(), np.unravel_index(excess.argmax(), excess.shape))
    return float(d.max())
