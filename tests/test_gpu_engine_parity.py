"""GPU parity: the CUDA engine (through the C ABI) against the CPU oracle.
Run on the B200 box:  python -m pytest tests -m gpu"""
import numpy as np
import pytest

import oracle_lib
import parity
import signals

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    return torch


def _engine(**kw):
    from gr_fosphor_b200.engine import Fosphor
    return Fosphor(**kw)


def _to_dev(torch, x):
    t = torch.from_numpy(np.ascontiguousarray(x).view(np.float32)).cuda()
    torch.cuda.synchronize()
    return t


# ---------------------------------------------------------------------------
# stage 1: the transform alone, every supported size
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("n", [512, 1024, 2048, 4096, 8192, 16384])
def test_fft_matches_double_dft(torch_cuda, n):
    torch = torch_cuda
    b = 16
    x = signals.noise_tones(n * b, n_fft=n, seed=100 + n, sigma=0.05)
    win = oracle_lib.default_window(n)
    eng = _engine(fft_len=n, n_bins=64, wf_rows=1024, window=win)
    xin = _to_dev(torch, x)
    out = torch.empty((b, n, 2), dtype=torch.float32, device="cuda")
    assert eng.debug_fft(xin.data_ptr(), b, n, out.data_ptr()) == 0
    eng.sync()
    got = out.cpu().numpy().view(np.complex64).reshape(b, n)
    prod = (x.reshape(b, n) * win[None, :]).astype(np.complex64)       # f32 products (fft.cl:416-417)
    ref = np.fft.fft(prod.astype(np.complex128), axis=1)
    err = np.abs(got - ref).max(axis=1) / np.abs(ref).max(axis=1)
    # max_k |X_gpu - X_ref| <= 1e-5 * max_k |X_ref| per spectrum
    assert err.max() <= 1e-5, err.max()
    eng.close()


@pytest.mark.parametrize("n", [2048, 4096])
def test_fft_radix64_plans_match_double_dft(torch_cuda, n, monkeypatch):
    """two-pass plans with a radix-64 register pass (FOSPHOR_B200_FFT_R64=1): the transform
    against the double DFT, and the whole path against the oracle."""
    monkeypatch.setenv("FOSPHOR_B200_FFT_R64", "1")
    test_fft_matches_double_dft(torch_cuda, n)
    b = 64
    x = signals.noise_tones(n * b, n_fft=n, seed=60 + n, sigma=0.02)
    cfg = dict(fft_len=n, n_bins=256, wf_rows=1024)
    eng, host, orc = _run_both(torch_cuda, cfg, [x])
    _check(cfg, host, orc, np.arange(b), b * n, single_call=True)
    eng.close()


def test_fft_hop_addressing(torch_cuda):
    """hop < N reads overlapping windows from the raw stream (overlap_cc_impl.cc:64-79)."""
    torch = torch_cuda
    n, ov, b = 1024, 4, 32
    raw = signals.noise_tones((b - 1) * (n // ov) + n, seed=5)
    eng = _engine()
    out_hop = torch.empty((b, n, 2), dtype=torch.float32, device="cuda")
    out_mat = torch.empty((b, n, 2), dtype=torch.float32, device="cuda")
    d_raw = _to_dev(torch, raw)
    d_mat = _to_dev(torch, signals.overlap_windows(raw, n, ov, b))
    eng.debug_fft(d_raw.data_ptr(), b, n // ov, out_hop.data_ptr())
    eng.debug_fft(d_mat.data_ptr(), b, n, out_mat.data_ptr())
    eng.sync()
    assert torch.equal(out_hop, out_mat)        # bit exact: same arithmetic, different addressing
    eng.close()


# ---------------------------------------------------------------------------
# whole path against the oracle
# ---------------------------------------------------------------------------
def _okw(cfg):
    return dict(fft_len=cfg.get("fft_len", 1024), n_bins=cfg.get("n_bins", 128),
                wf_rows=cfg.get("wf_rows", 1024), batch_mult=cfg.get("batch_mult", 16),
                batch_max=cfg.get("batch_max", 1024), t0r=cfg.get("t0r", 16.0),
                t0d=cfg.get("t0d", 1024.0), alpha=cfg.get("alpha", 0.002))


def _hrange(cfg):
    """(histo_scale, histo_ofs) of the default power range: scale * K, offset (cl.c:1087-1088)"""
    s, o = oracle_lib.power_range(cfg.get("fft_len", 1024), 0, 10)
    return np.float32(s) * np.float32(cfg.get("n_bins", 128)), o


def _run_both(torch, cfg, calls, via="device", twin=True):
    """calls: list of complex64 arrays (one per process call).  Runs them through the engine and
    the oracle (tier 2) and - twin=True - feeds the engine's OWN log-power rows, call by call, to a
    second oracle's display stage, whose histogram must equal the engine's bit for bit (tier 1,
    tests/parity.py).  Returns (engine, host arrays, oracle)."""
    eng = _engine(**cfg)
    okw = _okw(cfg)
    orc = oracle_lib.Oracle(**okw)
    tw = parity.DisplayTwin(**okw) if twin else None
    n, w = okw["fft_len"], okw["wf_rows"]
    keep, host, rc = [], None, 1
    for x in calls:
        assert orc.process(x) == 0
        pos, b = eng.waterfall_position, x.size // n
        if via == "device":
            d = _to_dev(torch, x)
            keep.append(d)
            assert eng.process_device(d.data_ptr(), b) == 0
        else:
            assert eng.process(x) == 0
        if tw is not None:
            rc, host = eng.finish()         # only copies: the state machine goes READY -> PENDING again
            assert rc == 1
            if b:
                tw.feed(host["waterfall"], pos, b)
    if tw is None or host is None:
        rc, host = eng.finish()
    assert rc == 1 and orc.finish() == 1
    assert eng.waterfall_position == orc.waterfall_position
    if tw is not None:
        tw.check(host["histogram"], host["spectrum"])
    return eng, host, orc


def _check(cfg, host, orc, rows, hits, single_call=False, max_skipped=0):
    hs, ho = _hrange(cfg)
    return parity.check_end_to_end(host, orc, rows, hits, hs, ho, single_call=single_call, max_skipped=max_skipped)


def _written_rows(calls, n, w):
    total = sum(x.size // n for x in calls)
    return np.arange(w) if total >= w else np.arange(total)


@pytest.mark.parametrize("n_bins", [128, 256])
@pytest.mark.parametrize("via", ["device", "host"])
def test_cfg1_burst(torch_cuda, n_bins, via):
    """BASELINE.json configs[0]: N=1024, one 64k burst (128 bins = reference, 256 = BASELINE)."""
    x = signals.cfg1_burst()
    cfg = dict(n_bins=n_bins)
    eng, host, orc = _run_both(torch_cuda, cfg, [x], via)
    # rows never written keep the first-use fill (cl.c:418-436)
    assert np.array_equal(host["waterfall"][64:], orc.waterfall[64:])
    _check(cfg, host, orc, np.arange(64), 64 * 1024, single_call=True)
    eng.close()


def test_call_sequence_state(torch_cuda):
    """state carried across calls of different batch sizes + ring wrap"""
    n = 1024
    sizes = (16, 64, 256, 1024, 32, 1024, 48)
    stream = signals.noise_tones(n * sum(sizes), seed=21)
    calls, pos = [], 0
    for b in sizes:
        calls.append(stream[pos:pos + b * n])
        pos += b * n
    eng, host, orc = _run_both(torch_cuda, dict(), calls)
    _check(dict(), host, orc, np.arange(1024), sum(sizes) * n)
    eng.close()


def test_cfg3_persistence_stress(torch_cuda):
    """BASELINE.json configs[2]: N=4096, 512 bins, overlap 8, fast decay (t0d=20)."""
    torch = torch_cuda
    n, k, ov, b = 4096, 512, 8, 256
    hop = n // ov
    raw = signals.burst_stress(n_fft=hop, n_spectra=2 * b + ov, seed=3, burst=64)   # bursts every 64 hops
    cfg = dict(fft_len=n, n_bins=k, wf_rows=1024, t0d=20.0)
    eng = _engine(**cfg)
    orc = oracle_lib.Oracle(fft_len=n, n_bins=k, wf_rows=1024, t0d=20.0)
    d_raw = _to_dev(torch, raw)
    for c in range(2):
        off = c * b * hop
        assert orc.process_hop(raw[off:], b, hop) == 0
        assert eng.process_device(d_raw.data_ptr() + 8 * off, b, hop) == 0
    rc, host = eng.finish()
    orc.finish()
    # tier 1: the engine's own rows through the oracle's display stage, bit-exact histogram
    tw = parity.DisplayTwin(fft_len=n, n_bins=k, wf_rows=1024, t0d=20.0)
    for c in range(2):
        tw.feed(host["waterfall"], c * b, b)
    tw.check(host["histogram"], host["spectrum"])
    _check(cfg, host, orc, np.arange(2 * b), 2 * b * n)
    eng.close()


def test_cfg4_large(torch_cuda):
    """BASELINE.json configs[3] shape on one GPU: N=16384, 1024 bins, B=1024 (one channel)."""
    n, k, b = 16384, 1024, 1024
    x = signals.noise_tones(n * b, n_fft=n, seed=10, sigma=0.02)
    cfg = dict(fft_len=n, n_bins=k, wf_rows=1024)
    eng, host, orc = _run_both(torch_cuda, cfg, [x])
    _check(cfg, host, orc, np.arange(1024), b * n, single_call=True)
    eng.close()


@pytest.mark.parametrize("n", [512, 2048, 8192])
def test_sweep_sizes(torch_cuda, n):
    """BASELINE.json configs[4] sizes not covered above, 256 bins."""
    b = 64
    x = signals.noise_tones(n * b, n_fft=n, seed=50 + n, sigma=0.02)
    cfg = dict(fft_len=n, n_bins=256, wf_rows=1024)
    eng, host, orc = _run_both(torch_cuda, cfg, [x])
    _check(cfg, host, orc, np.arange(b), b * n, single_call=True)
    eng.close()


def test_dc_contention_and_zeros(torch_cuda):
    """worst-case atomic contention (constant input: one bin per column) and the -inf path."""
    n, b = 1024, 256
    dc = np.full(n * b, 0.25 + 0.1j, np.complex64)
    eng, host, orc = _run_both(torch_cuda, dict(), [dc])
    rows_written = np.arange(b)
    # a constant under the periodic Hamming window is exactly three lines (bins 0, +-1): every other
    # bin holds f32 rounding noise - its waterfall / bins / live values are not comparable (those
    # columns are skipped and counted), but tier 1 above compared the engine's own rows exactly
    hs, ho = _hrange(dict())
    lines = np.array([0, 1, 1023])
    parity.check_waterfall(host["waterfall"][:, lines], orc.waterfall[:, lines], rows=rows_written)
    for key in ("histogram",):
        assert np.abs(host[key][:, lines] - orc.histogram[:, lines]).max() <= parity.HIST_TOL
    parity.check_spectrum(host["spectrum"], orc.spectrum, wf_ref=orc.waterfall[rows_written], max_skipped=1021)
    eng.close()

    zeros = np.zeros(n * 16, np.complex64)
    data = signals.noise_tones(n * 32, seed=11)
    eng, host, orc = _run_both(torch_cuda, dict(), [zeros, data])
    wf = host["waterfall"]
    assert np.all(np.isneginf(wf[:16])) and np.all(np.isneginf(orc.waterfall[:16]))
    _check(dict(), host, orc, np.arange(16, 48), 48 * n)
    eng.close()


def test_validation_and_state_machine(torch_cuda):
    """cl.c:881-886 / :970-994 return codes"""
    eng = _engine()
    rc, host = eng.finish()                 # BOOTING: clears, returns 1
    assert rc == 1
    scale, offset = oracle_lib.power_range(1024, 0, 10)
    assert np.all(host["waterfall"] == -offset) and np.all(host["spectrum"] == -offset)
    assert np.all(host["histogram"] == 0)
    assert eng.finish()[0] == 0             # READY: nothing new
    assert eng.process(np.zeros(1024 * 15, np.complex64)) == -22
    assert eng.process(np.zeros(1024 * 1040, np.complex64)) == -22
    assert eng.finish()[0] == 0
    assert eng.process(signals.noise_tones(1024 * 16, seed=1)) == 0
    assert eng.waterfall_position == 16
    assert eng.finish()[0] == 1
    eng.close()


def test_fallback_paths_odd_batches_and_unaligned_hop(torch_cuda):
    """batch_mult = 8 (batches that are not multiples of 16 rows -> plain count kernel,
    ragged 128-row blocks) and an odd hop (spectra not 16-byte aligned -> plain FFT
    kernel instead of the TMA streaming one), against the oracle."""
    torch = torch_cuda
    n = 1024
    sizes = (8, 24, 136, 1000, 16)
    stream = signals.noise_tones(n * sum(sizes), seed=91)
    calls, pos = [], 0
    for b in sizes:
        calls.append(stream[pos:pos + b * n])
        pos += b * n
    eng, host, orc = _run_both(torch, dict(batch_mult=8), calls)
    _check(dict(), host, orc, np.arange(1024), sum(sizes) * n)
    eng.close()

    hop, b = 255, 64
    raw = signals.noise_tones((b - 1) * hop + n, seed=92)
    eng = _engine()
    orc = oracle_lib.Oracle()
    d = _to_dev(torch, raw)
    assert eng.process_device(d.data_ptr(), b, hop) == 0
    assert orc.process_hop(raw, b, hop) == 0
    _, host = eng.finish()
    orc.finish()
    _check(dict(), host, orc, np.arange(b), b * n, single_call=True)
    eng.close()


def test_n512_pairs_with_odd_spectrum_counts(torch_cuda):
    """N = 512: a warp of the streaming FFT kernel transforms two spectra at a time; odd
    call sizes leave the last pair half empty and shift the pairing of the next call."""
    torch = torch_cuda
    n = 512
    sizes = (7, 33, 1, 128, 5)
    stream = signals.noise_tones(n * sum(sizes), n_fft=n, seed=93)
    calls, pos = [], 0
    for b in sizes:
        calls.append(stream[pos:pos + b * n])
        pos += b * n
    cfg = dict(fft_len=n, n_bins=128, batch_mult=1)
    eng, host, orc = _run_both(torch, cfg, calls)
    _check(cfg, host, orc, np.arange(sum(sizes)), sum(sizes) * n)
    eng.close()


def test_kernel_variants_are_bit_identical(torch_cuda, monkeypatch):
    """Transport variants must not change a single bit: TMA-prefetching FFT kernels
    (warp-level for N = 512/1024, CTA-level for 2048/4096/8192, half-staged for 16384) vs plain; fused
    accumulate kernel with 256/64/16-row TMA boxes vs plain loads;
    two-stream overlap vs one stream (several calls folded per launch).  The older
    split count/update kernels sum the live spectrum in another order: histogram and
    waterfall still bit-identical, live/max within the parity tolerance."""
    torch = torch_cuda
    names = ("FFT_VARIANT", "OVERLAP", "ACC", "ACC_BOX", "ACC_SUB", "COUNT_VARIANT", "OVERLAP_CHUNK", "ACC_SLIM",
             "ACC_GROUP")
    variants = (("2", "0", "1", "256", "64", "1", "16", "1", "1"),     # fused kernel, one stream
                ("0", "0", "1", "16", "16", "1", "16", "1", "2"),      # calls handed over in pairs
                ("2", "0", "1", "256", "64", "1", "16", "1", "4"),     # ... in fours (short last group)
                ("4", "0", "1", "64", "16", "1", "16", "1", "1"),      # grouped FFT kernel (8192 / 16384: warp-local late passes)
                ("1", "1", "1", "64", "64", "1", "2", "1", "0"),       # two streams, slim co-resident accumulate CTAs
                ("2", "1", "1", "256", "16", "1", "1", "1", "0"),
                ("2", "1", "1", "256", "64", "1", "1", "0", "2"),      # two streams, full-size accumulate CTAs
                ("3", "0", "1", "0", "16", "1", "16", "1", "4"),       # plain loads in the fused kernel
                ("2", "1", "1", "0", "16", "1", "1", "1", "1"),        # ... and in the slim one
                ("2", "1", "0", "64", "64", "1", "1", "1", "0"),       # split kernels, TMA-staged count
                ("0", "0", "0", "64", "64", "0", "16", "1", "0"))      # split kernels, plain count
    for n in (1024, 512, 2048, 4096, 8192, 16384):
        calls, b = (5, 1024) if n <= 1024 else (3, 256)
        x = signals.noise_tones(n * b * calls, n_fft=n, seed=77)
        d = _to_dev(torch, x)
        outs = []
        for var in variants:
            for k, v in zip(names, var):
                monkeypatch.setenv("FOSPHOR_B200_" + k, v)
            e = _engine(fft_len=n, n_bins=256, wf_rows=4096)
            assert e.process_device_multi(d.data_ptr(), calls, b) == 0
            _, h = e.finish()
            outs.append({k: v.copy() for k, v in h.items()})
            e.close()
        for var, o in zip(variants[1:], outs[1:]):
            for key in ("waterfall", "histogram"):
                assert np.array_equal(outs[0][key], o[key]), (n, key, var)
            if var[2] == "1":
                assert np.array_equal(outs[0]["spectrum"], o["spectrum"]), (n, "spectrum", var)
            else:
                # identical waterfall rows, another f32 summation order of the live spectrum
                parity.check_spectrum(o["spectrum"], outs[0]["spectrum"], tol0=parity.TWIN_SPEC_TOL)
        assert np.array_equal(outs[-1]["spectrum"], outs[-2]["spectrum"]), n


# ---------------------------------------------------------------------------
# size-independent properties at full size
# ---------------------------------------------------------------------------
def test_multi_call_launch_equals_single_calls(torch_cuda):
    """process_device_multi == the same calls one by one, bit for bit (cfg2 shape)."""
    torch = torch_cuda
    n, k, ov, b, calls = 1024, 256, 4, 1024, 6
    hop = n // ov
    raw = signals.noise_tones((calls * b - 1) * hop + n, seed=2)
    d_raw = _to_dev(torch, raw)
    a = _engine(n_bins=k, wf_rows=4096)
    c = _engine(n_bins=k, wf_rows=4096)
    assert a.process_device_multi(d_raw.data_ptr(), calls, b, hop) == 0
    for i in range(calls):
        assert c.process_device(d_raw.data_ptr() + 8 * i * b * hop, b, hop) == 0
    _, ha = a.finish()
    _, hc = c.finish()
    for key in ("waterfall", "histogram", "spectrum"):
        assert np.array_equal(ha[key], hc[key]), key
    assert a.waterfall_position == c.waterfall_position == (calls * b) % 4096
    # and host-fed raw stream == device path
    h = _engine(n_bins=k, wf_rows=4096)
    assert h.process_host_raw(raw, calls, b, hop) == 0
    _, hh = h.finish()
    for key in ("waterfall", "histogram", "spectrum"):
        assert np.array_equal(ha[key], hh[key]), key
    for e in (a, c, h):
        e.close()


@pytest.mark.parametrize("b", [16, 64, 256, 528])
def test_small_batches_folded_launches_vs_oracle(torch_cuda, b):
    """Many small calls folded into few launches (groups of 4 calls per counter -> updater
    hand-over for B <= 256, short last group, runs shorter than 64 rows, ring wrap between
    launches) against the oracle fed the same calls one by one, and against the engine fed one
    by one (bit exact)."""
    torch = torch_cuda
    n, k, calls = 2048, 128, 11
    x = signals.noise_tones(n * b * calls, n_fft=n, seed=300 + b, sigma=0.03)
    d = _to_dev(torch, x)
    rows = {16: 128, 64: 512, 256: 2048, 528: 4096}[b]           # 11 calls wrap every ring
    bmax = 1024 if b > 256 else b
    cfg = dict(fft_len=n, n_bins=k, wf_rows=rows, t0d=64.0, batch_max=bmax)
    eng = _engine(**cfg)
    one = _engine(**cfg)
    orc = oracle_lib.Oracle(fft_len=n, n_bins=k, wf_rows=rows, t0d=64.0, batch_max=bmax)
    assert eng.process_device_multi(d.data_ptr(), calls, b) == 0
    for c in range(calls):
        assert one.process_device(d.data_ptr() + 8 * c * b * n, b) == 0
        assert orc.process(x[c * b * n:(c + 1) * b * n]) == 0
    _, host = eng.finish()
    _, host1 = one.finish()
    orc.finish()
    for key in ("waterfall", "histogram", "spectrum"):
        assert np.array_equal(host[key], host1[key]), key
    assert eng.waterfall_position == orc.waterfall_position
    assert calls * b > rows
    _check(cfg, host, orc, np.arange(rows), calls * b * n)
    for e in (eng, one):
        e.close()


def test_automatic_two_stream_schedule_is_bit_identical(torch_cuda, monkeypatch):
    """A ring of four chunks of >= 32 M samples switches the engine to its two-stream schedule
    (accumulate of chunk c beside the FFT of chunk c+1, streams joined lazily: three process calls
    in a row without a finish).  Results must equal the one-stream schedule bit for bit."""
    torch = torch_cuda
    n, k, b, calls, hop, rows = 1024, 256, 1024, 96, 256, 131072
    raw = signals.noise_tones((calls * b - 1) * hop + n, seed=404)
    d_raw = _to_dev(torch, raw)
    outs = []
    for mode in (None, "0"):
        if mode is None:
            monkeypatch.delenv("FOSPHOR_B200_OVERLAP", raising=False)
        else:
            monkeypatch.setenv("FOSPHOR_B200_OVERLAP", mode)
        e = _engine(n_bins=k, wf_rows=rows)
        for rep in range(3):                      # the ring wraps; no join between the calls
            assert e.process_device_multi(d_raw.data_ptr(), calls, b, hop) == 0
        assert (e.two_stream_chunks > 0) == (mode is None)
        _, h = e.finish()
        outs.append({key: v.copy() for key, v in h.items()})
        assert e.waterfall_position == (3 * calls * b) % rows
        e.close()
    for key in ("waterfall", "histogram", "spectrum"):
        assert np.array_equal(outs[0][key], outs[1][key]), key


@pytest.mark.parametrize("b", [8192, 32768])
def test_maximum_batch_sizes(torch_cuda, b):
    """Batches beyond the shared-memory (d, e) table (B > 4096: table read from global memory)
    up to the largest batch the ABI accepts (32768), one call, N = 512."""
    n, k = 512, 128
    x = signals.noise_tones(n * b, n_fft=n, seed=500 + b, sigma=0.05)
    cfg = dict(fft_len=n, n_bins=k, wf_rows=b, batch_max=b)
    eng, host, orc = _run_both(torch_cuda, cfg, [x])
    _check(cfg, host, orc, np.arange(b), b * n, single_call=True)
    eng.close()


def test_histogram_mass_property(torch_cuda):
    """From a zero histogram one call deposits, per column, exactly the mass the
    closed form predicts from hit counts summing to B: sum_bins d(hc)*(1-e(hc)).
    Checked against the oracle's hit counts at the full cfg2 batch size."""
    n, k, b = 1024, 256, 1024
    x = signals.noise_tones(n * b, seed=33)
    eng, host, orc = _run_both(torch_cuda, dict(n_bins=k), [x])
    hits = orc.last_hits
    rows_written = np.arange(b)
    assert np.all(hits.sum(axis=0) == b)
    _check(dict(n_bins=k), host, orc, rows_written, b * n, single_call=True)
    # live IIR linearity in the carry: second identical call moves live towards the same mean
    eng.close()
