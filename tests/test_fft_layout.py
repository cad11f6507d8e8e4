"""Host-side checks of the FFT kernels' shared-memory exchange layout (gr-fosphor_b200/csrc/
fft_power.cuh: FftPlan::PADSHIFT, pad_idx, pad_step), restated in Python:

* pad_step: pad(base + t*stride) == pad(base) + const(t) for every access pattern of every plan
  (the kernels use the constants as immediates);
* the padding makes every 64-bit access of a half-warp (16 consecutive threads, one wavefront)
  hit 16 different bank pairs - no replays - for every pass of every plan.
"""
import pytest

PLANS = [(512, 16, 32, 2), (1024, 32, 32, 2), (2048, 8, 16, 3), (4096, 16, 16, 3),
         (8192, 8, 32, 3), (16384, 16, 32, 3), (2048, 32, 64, 2), (4096, 64, 64, 2)]


def _layout(n, r0, r1):
    shift = max(4, r0.bit_length() - 1)            # FftPlan::PADSHIFT
    two = (r0, r1) == (8, 32)                      # FftPlan::PADSHIFT2 = 8, PAD2 = 8 (N = 8192)
    pad = lambda a: a + (a >> shift) + (8 * (a >> 8) if two else 0)              # pad_idx
    step = lambda t, stride: t * stride + ((t * stride) >> shift) + (8 * ((t * stride) >> 8) if two else 0)   # pad_step
    return pad, step, n // r0, n // r1


def _patterns(n, r0, r1, npass):
    """(name, threads, index(i, t), base(i), stride, t range)"""
    pad, step, nb0, nb1 = _layout(n, r0, r1)
    pats = [("pass0 store", nb0, lambda i, t: i * r0 + t, lambda i: i * r0, 1, r0),
            ("pass1 load", nb1, lambda i, t: i + t * nb1, lambda i: i, nb1, r1)]
    if npass == 3:
        j = lambda i: (i - (i & (r0 - 1))) * r1 + (i & (r0 - 1))
        pats.append(("pass1 store", nb1, lambda i, t: j(i) + t * r0, j, r0, r1))
        pats.append(("pass2 load", nb1, lambda i, t: i + t * nb1, lambda i: i, nb1, r1))
    return pats


@pytest.mark.parametrize("plan", PLANS)
def test_pad_step_identity(plan):
    n, r0, r1, npass = plan
    pad, step, _, _ = _layout(n, r0, r1)
    for name, threads, index, base, stride, tn in _patterns(*plan):
        for i in range(threads):
            for t in range(tn):
                assert pad(index(i, t)) == pad(base(i)) + step(t, stride), (plan, name, i, t)


@pytest.mark.parametrize("plan", PLANS)
def test_exchange_accesses_are_bank_conflict_free(plan):
    n, r0, r1, npass = plan
    pad, _, _, _ = _layout(n, r0, r1)
    for name, threads, index, _, _, tn in _patterns(*plan):
        lanes = min(16, threads)
        for i0 in range(0, threads, lanes):
            for t in range(tn):
                pairs = {(pad(index(i, t)) * 2 % 32) // 2 for i in range(i0, i0 + lanes)}
                assert len(pairs) == lanes, (plan, name, i0, t)
