"""Tolerance helpers shared by the parity tests.

The path is floating point and the reference itself is not bit-reproducible
across OpenCL runtimes (native_sin/cos/powr, fft.cl:66-67,
display.cl:150,210,242-245), so parity is checked in two tiers:

Tier 1 - stage-wise, EXACT where the arithmetic is exact.  The display stage
  (bin map, hit counts, rise/decay: display.cl:160-254) is integer / exactly
  rounded f32 work once the log-power rows are fixed.  `DisplayTwin` feeds the
  rows the CUDA engine wrote into ITS waterfall to the oracle's display stage
  (fosphor_oracle_process_pwr); the engine's histogram must then equal the
  oracle's BIT FOR BIT and live / max-hold to 2e-5 (different f32 summation
  order only).  No flip budget: a hit moved by any number of bins, a corrupted
  tile or a dropped row fails this.

Tier 2 - end to end against the oracle / the reference's golden outputs
  (different FFT arithmetic upstream, so rounding-boundary bin flips exist):
  waterfall   |d pwr| <= 1e-4 log10 units, or |d mag| <= 1e-5 * max(mag) per row
              (bins far below the row maximum are ill-conditioned in log).
              With histo_scale * 1e-4 << 1 this already implies |d bin| <= 1 for
              every well-conditioned hit: ZERO hits move by two bins or more.
  flips       `count_flips` maps both waterfalls to bins exactly like the kernel and
              counts the hits whose bin differs (and asserts none differs by >= 2 on
              well-conditioned cells, and at most FLIP_FRAC of the hits flip).
  histogram   |d hv| <= 2e-3 except for cells touched by a flip.  Budget for such
              cells: 5 x the OBSERVED flip count (+8), or 1e-3 of the hits in play
              when no waterfall is at hand (measured oracle-vs-reference: 2.4e-4).
              Single calls from the cleared state additionally require every
              out-of-tolerance cell to have a partner whose difference has the
              opposite sign in an ADJACENT bin of the same column (a +-1-bin flip
              moves mass between neighbours; later calls contract the two cells at
              different rates, so the rule is only sound for the first call).
  live / max  |d| <= 1e-4 + the error an f32 FFT puts on the rows that fed the column
              (spectrum_tolerance); columns whose bound exceeds 0.02 are ill-conditioned
              and skipped - how many may be is bounded by the caller (default: none).
"""
import numpy as np

PWR_TOL = 1e-4
MAG_REL_TOL = 1e-5
HIST_TOL = 2e-3
SPEC_TOL = 1e-4
FLIP_FRAC = 0.005           # at most 0.5 % of the hits may flip by one bin
FLIP_BUDGET_X = 5           # histogram cells out of tolerance <= 5 x observed flips (+ slack)
FLIP_BUDGET_SLACK = 8
NOFLIP_BUDGET_FRAC = 1e-3   # ... or this fraction of the hits when flips cannot be observed
TWIN_SPEC_TOL = 2e-5        # tier 1: live / max-hold on identical rows (summation order only)


def _eq_nonfinite(a, b):
    return (~np.isfinite(a)) & (~np.isfinite(b)) & ((a == b) | (np.isnan(a) & np.isnan(b)))


def check_waterfall(got, ref, rows=None):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    if rows is not None:
        got, ref = got[rows], ref[rows]
    same = _eq_nonfinite(got, ref)
    assert np.all(np.isfinite(got) | same), "non-finite waterfall values differ"
    with np.errstate(invalid="ignore"):          # inf - inf where both are the same infinity: masked by `same`
        d = np.where(same, 0.0, np.abs(got - ref))
    mag_g, mag_r = 10.0 ** np.where(same, 0, got), 10.0 ** np.where(same, 0, ref)
    rowmax = np.maximum(mag_r.max(axis=-1, keepdims=True), 1e-300)
    ok = (d <= PWR_TOL) | (np.abs(mag_g - mag_r) <= MAG_REL_TOL * rowmax)
    assert ok.all(), "waterfall: %d cells out of tolerance, worst |dpwr|=%g" % ((~ok).sum(), d[~ok].max())
    return float(d.max())


def bins_from_rows(rows, hscale, hofs, n_bins):
    """display.cl:160-168 in exact f32: bin = (int)round(hscale * (pwr + hofs)), half away
    from zero, clamped to [0, K-1]; NaN / -inf -> 0, +inf -> K-1 (DESIGN.md fixed rule)."""
    f = np.float32
    with np.errstate(invalid="ignore", over="ignore"):
        x = (f(hscale) * (np.asarray(rows, np.float32) + f(hofs))).astype(np.float32)
        t = np.trunc(x)
        r = t + ((x - t) >= f(0.5))             # x - trunc(x) is exact in f32
        r = np.where(x > 0, r, 0.0)             # negatives, NaN -> 0 (lower clamp)
        r = np.minimum(r, float(n_bins - 1))    # +inf -> K-1
    return np.nan_to_num(r, nan=0.0).astype(np.int64)


def count_flips(wf_got, wf_ref, hscale, hofs, n_bins, rows=None):
    """Hits whose bin differs between the two waterfalls.  Returns the flip count; asserts that
    no well-conditioned hit (|d pwr| <= PWR_TOL) moved by two bins or more and that at most
    FLIP_FRAC of the hits flipped."""
    g = np.asarray(wf_got, np.float32)
    r = np.asarray(wf_ref, np.float32)
    if rows is not None:
        g, r = g[rows], r[rows]
    bg, br = bins_from_rows(g, hscale, hofs, n_bins), bins_from_rows(r, hscale, hofs, n_bins)
    db = np.abs(bg - br)
    with np.errstate(invalid="ignore"):
        well = np.abs(g.astype(np.float64) - r.astype(np.float64)) <= PWR_TOL
    assert not np.any(well & (db >= 2)), "%d well-conditioned hits moved by >= 2 bins" % int((well & (db >= 2)).sum())
    flips = int((db != 0).sum())
    assert flips <= FLIP_FRAC * db.size + 2, "%d of %d hits flipped (> %.1f %%)" % (flips, db.size, 100 * FLIP_FRAC)
    return flips


def check_histogram(got, ref, hits_in_play, flips=None, visible_hits=None, single_call=False, t0r=16.0):
    """flips: observed bin flips (count_flips) on `visible_hits` hits (default: all of
    hits_in_play); the out-of-tolerance cell budget is FLIP_BUDGET_X times that, scaled to the
    hits in play.  single_call: first call after the clears - strict +-1-bin adjacency rule."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    diff = got - ref
    d = np.abs(diff)
    badm = d > HIST_TOL
    bad = int(badm.sum())
    if flips is not None:
        scale = 1.0 if not visible_hits else max(1.0, hits_in_play / float(visible_hits))
        # flips seen on a SAMPLE of the rows (fixtures keep 32 of 1024) are extrapolated with +2 for
        # the sampling noise of a small count
        budget = int(FLIP_BUDGET_X * (flips + (2 if scale > 1.0 else 0)) * scale) + FLIP_BUDGET_SLACK
    else:
        budget = int(NOFLIP_BUDGET_FRAC * hits_in_play) + FLIP_BUDGET_SLACK
    assert bad <= budget, "histogram: %d cells differ by > %g (budget %d from %s flips), worst %g" % (
        bad, HIST_TOL, budget, flips, d.max())
    if single_call and bad:
        # the partner's difference may be far below HIST_TOL: one hit more or less barely moves a
        # cell that collected hundreds (hv saturates), while it lifts an empty neighbour by ~0.04
        sign = np.sign(diff) * (d > 1e-7)
        up = np.zeros_like(sign)
        dn = np.zeros_like(sign)
        up[:-1] = sign[1:]                      # bin k+1
        dn[1:] = sign[:-1]                      # bin k-1
        paired = badm & ((up == -sign) | (dn == -sign))
        # two flips in one column can chain (k-1 -> k and k -> k+1 in different rows: bin k keeps its
        # count, the ends differ): accept an out-of-tolerance partner of the opposite sign two bins away
        bsign = np.sign(diff) * badm
        up2 = np.zeros_like(sign)
        dn2 = np.zeros_like(sign)
        up2[:-2] = bsign[2:]
        dn2[2:] = bsign[:-2]
        paired |= badm & ((up2 == -bsign) | (dn2 == -bsign))
        lone = badm & ~paired
        assert not lone.any(), "histogram: %d out-of-tolerance cells without an opposite-sign neighbour bin " \
            "(not a +-1-bin flip), first at %s" % (int(lone.sum()), np.argwhere(lone)[0])
    # a flipped hit moves mass to the neighbouring bin only: column sums stay close.  Not equal: the rise
    # is (1 - hv) / t0r per hit, so a hit that leaves a saturated cell for an empty one adds up to 1 / t0r
    cs = np.abs(got.sum(axis=0) - ref.sum(axis=0))
    assert cs.max() <= max(0.05, 1.5 / t0r) + 1e-3 * ref.sum(axis=0).max(), "histogram column mass differs: %g" % cs.max()
    return bad, float(d.max())


FFT_REL = 1e-6      # error of an f32 FFT relative to the row maximum (ours measures ~3e-7)
SPEC_SKIP_TOL = 0.02    # a column whose error bound exceeds this is ill-conditioned: not compared, but counted
LIVE_ALPHA = 0.002  # cl.c:716
MH_MIX = 0.001      # display.cl:303


def cell_error_bound(wf_rows_ref):
    """Per waterfall cell: how far two correct f32 FFTs may disagree in log10 units.  A bin
    `depth` (log10) below the strongest line of its spectrum carries an absolute rounding error of
    about FFT_REL * rowmax in ANY f32 FFT (the reference's included): log10(1 + FFT_REL * 10**depth)
    in log10 units - FFT_REL * 10**depth / ln(10) for well-conditioned bins, depth - 6 for bins
    below the rounding floor.  -inf cells are compared exactly (bound 0)."""
    wf = np.asarray(wf_rows_ref, np.float64)
    fin = np.isfinite(wf)
    rowmax = np.max(np.where(fin, wf, -np.inf), axis=1, keepdims=True)
    rowmax = np.where(np.isfinite(rowmax), rowmax, 0.0)
    with np.errstate(invalid="ignore", over="ignore"):
        e = np.log10(1.0 + FFT_REL * 10.0 ** np.minimum(np.where(fin, rowmax - wf, 0.0), 30.0))
    return np.where(fin, e, 0.0)


def spectrum_tolerance(wf_rows_ref, tol0=SPEC_TOL):
    """(tol_live, tol_max) per DISPLAY-order column from the cell bounds of the rows that fed them.
    live = alpha * sum_s w_s pwr_s with w_s <= 1 and alpha * sum w_s <= 1: its error is at most
    min(1, 2 alpha R) times the mean cell bound of the R rows.  max-hold = max_s pwr_s (plus
    MH_MIX * live): |max(a) - max(b)| <= max_s(b_s + e_s) - max_s(b_s) - deep cells of a column
    whose maximum is strong do not matter."""
    wf = np.asarray(wf_rows_ref, np.float64)
    e = cell_error_bound(wf)
    r = wf.shape[0]
    tol_live = tol0 + min(1.0, 2.0 * LIVE_ALPHA * r) * e.mean(axis=0)
    fin = np.isfinite(wf)
    hi = np.max(np.where(fin, wf + e, -np.inf), axis=0)
    lo = np.max(np.where(fin, wf, -np.inf), axis=0)
    with np.errstate(invalid="ignore"):
        dmax = np.where(np.isfinite(hi) & np.isfinite(lo), hi - lo, 0.0)
    tol_max = tol0 + dmax + MH_MIX * tol_live
    n = wf.shape[1]
    perm = np.arange(n) ^ (n // 2)              # display.cl:201
    return tol_live[perm], tol_max[perm]


def check_spectrum(got, ref, cols=None, wf_ref=None, max_skipped=0, tol0=SPEC_TOL):
    """live / max-hold against per-column tolerances derived from the waterfall rows that fed them
    (spectrum_tolerance; flat tol0 without wf_ref).  Columns whose bound exceeds SPEC_SKIP_TOL hold
    only rounding noise (e.g. the off-peak bins of a pure tone under a rectangular window): they are
    not compared, and at most `max_skipped` columns may be (default: none - the caller must know
    that its signal has such columns).  Returns the worst deviation on the compared columns."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    n = got.shape[1]
    tol = np.stack([np.full(n, tol0), np.full(n, tol0)])
    if wf_ref is not None:
        tl, tm = spectrum_tolerance(wf_ref, tol0)
        tol = np.stack([tl, tm])
    # column-wide: max-hold mixes the live value in at every call (display.cl:303), so a column whose
    # live trace is ill-conditioned has no trustworthy max-hold either
    keep = np.broadcast_to((tol <= SPEC_SKIP_TOL).all(axis=0), tol.shape)
    skipped = int((~keep).any(axis=0).sum())
    assert skipped <= max_skipped, "spectrum: %d ill-conditioned columns would be skipped (allowed %d)" % (
        skipped, max_skipped)
    if cols is not None:
        keep = keep & np.asarray(cols)[None, :]
    same = _eq_nonfinite(got, ref)
    with np.errstate(invalid="ignore"):
        d = np.where(same, 0.0, np.abs(got - ref))
    d = np.nan_to_num(d, nan=np.inf)
    assert np.all(d[:, :, 0] <= 1e-6), "x coordinates differ"
    excess = np.where(keep, d[:, :, 1] - tol, -1.0)
    assert excess.max() <= 0, "spectrum: |d|=%g exceeds its tolerance %g at (trace, column) %s" % (
        d[:, :, 1][np.unravel_index(excess.argmax(), excess.shape)],
        tol[np.unravel_index(excess.argmax(), excess.shape)], np.unravel_index(excess.argmax(), excess.shape))
    return float(np.where(keep, d[:, :, 1], 0.0).max())


class DisplayTwin:
    """Tier 1: the oracle's display stage run on the ENGINE's own log-power rows.

    Feed it, call by call, the rows the engine wrote into its waterfall ring; `check`
    then demands a bit-identical histogram and live / max-hold within TWIN_SPEC_TOL."""

    def __init__(self, **oracle_kw):
        import oracle_lib
        self.o = oracle_lib.Oracle(**oracle_kw)

    def feed(self, waterfall, pos, batch):
        """rows [pos, pos + batch) (mod ring) of a waterfall snapshot = one process call"""
        w = waterfall.shape[0]
        rows = waterfall[(pos + np.arange(batch)) % w]
        assert self.o.process_pwr(rows) == 0

    def check(self, histogram, spectrum):
        self.o.finish()
        ref_h = self.o.histogram
        neq = histogram != ref_h
        assert not neq.any(), "display stage on identical rows: %d histogram cells differ (worst %g, first %s)" % (
            int(neq.sum()), np.abs(histogram.astype(np.float64) - ref_h).max(), np.argwhere(neq)[0])
        ref_s = self.o.spectrum
        g = np.asarray(spectrum, np.float64)
        same = _eq_nonfinite(g, ref_s.astype(np.float64))
        with np.errstate(invalid="ignore"):
            d = np.where(same, 0.0, np.abs(g - ref_s))
        d = np.nan_to_num(d, nan=np.inf)
        assert d.max() <= TWIN_SPEC_TOL, "display stage on identical rows: live/max differ by %g" % d.max()
        return float(d.max())


def check_end_to_end(host, ref, rows, hits_in_play, hscale, hofs, single_call=False, max_skipped=0, t0r=16.0):
    """Tier 2 in one call: `host` / `ref` are dicts (or objects) with waterfall, histogram,
    spectrum; `rows` the ring rows that hold data of the calls checked; hscale = scale * n_bins.
    Returns {"flips", "bad_cells", "hits"}."""
    def g(o, k):
        return o[k] if isinstance(o, dict) else getattr(o, k)
    wf_h, wf_r = g(host, "waterfall"), g(ref, "waterfall")
    n_bins = g(ref, "histogram").shape[0]
    check_waterfall(wf_h, wf_r, rows=rows)
    flips = count_flips(wf_h, wf_r, hscale, hofs, n_bins, rows=rows)
    bad, _ = check_histogram(g(host, "histogram"), g(ref, "histogram"), hits_in_play, flips=flips,
                             visible_hits=len(rows) * wf_r.shape[1], single_call=single_call, t0r=t0r)
    check_spectrum(g(host, "spectrum"), g(ref, "spectrum"), wf_ref=wf_r[rows], max_skipped=max_skipped)
    return {"flips": flips, "bad_cells": bad, "hits": hits_in_play}
