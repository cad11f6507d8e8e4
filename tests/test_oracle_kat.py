"""Analytic known-answer tests pinning the CPU oracle.
All values follow from the reference sources cited in oracle/fosphor_oracle.h
with the default power range db_ref=0, db_per_div=10: scale=0.2,
offset=1.9896998, histo_scale=25.6 (fosphor.c:138-151, cl.c:1087)."""
import numpy as np
import pytest

import oracle_lib
import signals

N = 1024
RECT = np.ones(N, np.float32)


def test_default_window_and_range():
    w = oracle_lib.default_window(N)
    n = np.arange(N)
    ref = (0.54 - 0.46 * np.cos(2 * 3.141592 * n / N)) * 1.855       # fosphor.c:117, truncated pi
    assert np.abs(w - ref).max() < 2e-6
    scale, offset = oracle_lib.power_range(N, 0, 10)
    assert scale == np.float32(0.2)
    assert abs(offset - (-(np.log10(1024.0) - 5.0))) < 1e-6           # 1.9896998
    s2, o2 = oracle_lib.power_range(N, -10, 5)
    assert abs(s2 - 0.4) < 1e-7 and abs(o2 - (-(np.log10(1024.0) - 3.0))) < 1e-6


@pytest.mark.parametrize("amp,k,expect_bin,expect_pwr", [
    (1.0, 100, 127, 3.0103),      # 25.6*(3.0103+1.9897)=128 -> clamp 127 (display.cl:161-165)
    (0.1, 300, 102, 2.0103),      # 102.4 -> 102
    (0.01, 700, 77, 1.0103),      # 76.8 -> 77
])
def test_tone_bins(amp, k, expect_bin, expect_pwr):
    o = oracle_lib.Oracle(window=RECT)
    assert o.process(signals.tone(N, 16, k, amp)) == 0
    wf = o.waterfall
    assert np.allclose(wf[:16, k], expect_pwr, atol=2e-5)
    hits = o.last_hits
    assert hits[expect_bin, k] == 16 and hits[:, k].sum() == 16
    assert np.all(hits.sum(axis=0) == 16)


def test_impulse_every_column_bin_51():
    o = oracle_lib.Oracle(window=RECT)
    o.process(signals.impulse(N, 16))
    assert np.allclose(o.waterfall[:16], 0.0, atol=1e-6)              # |X| = 1 everywhere
    hits = o.last_hits
    assert np.all(hits[51] == 16) and hits.sum() == 16 * N            # 25.6*1.9897 = 50.94 -> 51


@pytest.mark.parametrize("b,rise,decay,carry", [
    (16, 0.6398172, 0.9844889, 0.9684759),
    (64, 0.9698086, 0.9393844, 0.8797421),
    (1024, 0.9846154, 0.3676997, 0.1287318),
])
def test_rise_decay_and_live_carry(b, rise, decay, carry):
    """display.cl:241-247: all B spectra hit one bin from hv=0 -> rise value;
    then a call that leaves that bin untouched multiplies by (1-1/1024)^B."""
    o = oracle_lib.Oracle(window=RECT)
    o.process(signals.tone(N, b, 100, 0.1))       # bin 102 in column 100
    h1 = o.histogram
    assert abs(h1[102, 100] - rise) < 2e-6
    _, offset = oracle_lib.power_range(N, 0, 10)
    live1 = o.spectrum[0, :, 1]
    o.process(signals.tone(N, b, 100, 0.01))      # bin 77 now; bin 102 only decays
    h2 = o.histogram
    assert abs(h2[102, 100] / h1[102, 100] - decay) < 2e-6
    assert abs(h2[77, 100] - rise) < 2e-6
    # live IIR (display.cl:210): y' = y*(1-a)^B + a*S with S = pwr * sum_s (1-a)^(B-1-s)
    i = 100 ^ 512
    geo = (1 - 0.998 ** b) / 0.002
    y1 = -float(offset) * carry + 0.002 * 2.0103 * geo
    assert abs(live1[i] - y1) < 2e-4
    y2 = y1 * carry + 0.002 * 1.0103 * geo
    assert abs(o.spectrum[0, i, 1] - y2) < 2e-4


def test_small_residue_is_never_decayed():
    """display.cl:237-238: cells with hv <= 0.01 and no hit are skipped."""
    o = oracle_lib.Oracle(window=RECT, t0d=4.0)   # fast decay to get below 0.01 quickly
    o.process(signals.tone(N, 16, 100, 0.1))
    for _ in range(3):
        o.process(signals.tone(N, 16, 100, 0.01))
    v = o.histogram[102, 100]
    assert 0.0 < v <= 0.01
    o.process(signals.tone(N, 16, 100, 0.01))
    assert o.histogram[102, 100] == v


def test_waterfall_ring_and_position():
    o = oracle_lib.Oracle()
    x = signals.noise_tones(N * 128, seed=4)
    assert o.process(x[:64 * N]) == 0 and o.waterfall_position == 64
    assert o.process(x[64 * N:]) == 0 and o.waterfall_position == 128
    single = oracle_lib.Oracle()
    single.process(x[:64 * N])
    assert np.array_equal(o.waterfall[:64], single.waterfall[:64])
    _, offset = oracle_lib.power_range(N, 0, 10)
    assert np.all(o.waterfall[128:] == -offset)                       # untouched rows keep the fill
    for _ in range(7):                                                # (pos + B) & 1023, cl.c:954
        o.process(x)
    assert o.waterfall_position == (128 + 7 * 128) % 1024


def test_validation_and_state_machine():
    o = oracle_lib.Oracle()
    _, offset = oracle_lib.power_range(N, 0, 10)
    assert o.finish() == 1                                            # BOOTING: clears (cl.c:982-994)
    assert np.all(o.spectrum == -offset) and np.all(o.waterfall == -offset) and np.all(o.histogram == 0)
    assert o.finish() == 0                                            # READY
    assert o.process(np.zeros(N * 15, np.complex64)) == -22           # cl.c:881-883
    assert o.process(np.zeros(N * 1040, np.complex64)) == -22         # cl.c:885-886
    assert o.finish() == 0
    assert o.process(signals.noise_tones(N * 16, seed=1)) == 0
    assert o.finish() == 1 and o.finish() == 0


def test_max_hold_per_call_decay():
    """display.cl:303: m = 0.999*m + 0.001*live_new once per call, then max with the batch max."""
    o = oracle_lib.Oracle(window=RECT)
    _, offset = oracle_lib.power_range(N, 0, 10)
    o.process(signals.tone(N, 16, 100, 1.0))
    i = 100 ^ 512
    assert abs(o.spectrum[1, i, 1] - 3.0103) < 2e-5                    # batch max wins
    live = o.spectrum[0, i, 1]
    o.process(signals.tone(N, 16, 100, 0.001))                        # pwr 0.0103 < held max
    live2 = o.spectrum[0, i, 1]
    assert abs(o.spectrum[1, i, 1] - (0.999 * 3.0103 + 0.001 * live2)) < 1e-5
    # x coordinates, display.cl:209,293
    assert np.allclose(o.spectrum[0, :, 0], np.arange(N) / 512.0 - 1.0)
    assert np.allclose(o.spectrum[1, :, 0], np.arange(N) / 512.0 - 1.0)


def test_zero_input_goes_to_minus_inf_and_bin0():
    o = oracle_lib.Oracle()
    o.process(np.zeros(N * 16, np.complex64))
    assert np.all(np.isneginf(o.waterfall[:16]))
    assert np.all(o.last_hits[0] == 16)
    assert np.all(np.isneginf(o.spectrum[0, :, 1]))
    o.process(signals.noise_tones(N * 16, seed=3))                    # !isfinite fallbacks (display.cl:206-207,290-291)
    assert np.all(np.isfinite(o.spectrum[:, :, 1]))


@pytest.mark.parametrize("n", [512, 2048, 4096, 16384])
def test_oracle_fft_sizes(n):
    x = signals.noise_tones(n * 16, n_fft=n, seed=n)
    o = oracle_lib.Oracle(fft_len=n, n_bins=64)
    assert o.process(x) == 0
    w = oracle_lib.default_window(n)
    prod = (x.reshape(16, n) * w[None, :]).astype(np.complex64).astype(np.complex128)
    ref = np.fft.fft(prod, axis=1)
    assert np.abs(o.last_fft - ref).max() / np.abs(ref).max() < 2e-7


def test_f32_fft_variant_close_to_double():
    x = signals.cfg1_burst()
    a, b = oracle_lib.Oracle(), oracle_lib.Oracle(fft_f32=True)
    a.process(x), b.process(x)
    ref = np.abs(a.last_fft).max()
    assert np.abs(a.last_fft - b.last_fft).max() / ref < 1e-5


def test_hop_addressing_equals_materialised_overlap():
    """lib/overlap_cc_impl.cc:64-79"""
    n, ov, b = 1024, 4, 64
    raw = signals.noise_tones((b - 1) * (n // ov) + n, seed=6)
    a, c = oracle_lib.Oracle(), oracle_lib.Oracle()
    assert a.process_hop(raw, b, n // ov) == 0
    assert c.process(signals.overlap_windows(raw, n, ov, b)) == 0
    assert np.array_equal(a.waterfall, c.waterfall) and np.array_equal(a.histogram, c.histogram)
