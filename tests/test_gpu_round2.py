"""GPU tests of the engine features that carry the drop-in and the multi-GPU runs:
fold depth decoupled from the waterfall size (scratch ring + publish), new-rows read-back,
pageable host feed (copy threads, optional host registration), side-stream max-hold export,
several engines on several devices in one process, the K = 1024 limits, and BASELINE
configs[3] (one channel per GPU + NCCL max reduce) as a committed multi-rank test.
Run on the B200 box:  python -m pytest tests -m gpu   (multi-GPU tests skip on one GPU)"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib
import parity
import signals

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def torch_cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    return torch


def _engine(**kw):
    from gr_fosphor_b200.engine import Fosphor
    return Fosphor(**kw)


def _to_dev(torch, x, device="cuda"):
    t = torch.from_numpy(np.ascontiguousarray(x).view(np.float32)).to(device)
    torch.cuda.synchronize()
    return t


def test_fold_depth_is_independent_of_waterfall_rows(torch_cuda):
    """W = 1024 (the reference geometry, cl.c:430-432) with 24 calls of 1024 spectra in ONE
    process_device_multi: the engine grows a scratch ring, folds the calls into few launches and
    publishes the last 1024 rows into the waterfall - bit-identical to an engine that may not
    (scratch_rows = -1: one launch pair per call) and to calls made one by one."""
    torch = torch_cuda
    n, k, b, calls, hop = 1024, 256, 1024, 24, 256
    raw = signals.noise_tones((calls * b - 1) * hop + n, seed=77)
    d = _to_dev(torch, raw)
    deep = _engine(n_bins=k, wf_rows=1024)                     # automatic scratch ring
    flat = _engine(n_bins=k, wf_rows=1024, scratch_rows=-1)    # the waterfall is the ring
    one = _engine(n_bins=k, wf_rows=1024, scratch_rows=-1)
    l0 = deep.launch_count
    assert deep.process_device_multi(d.data_ptr(), calls, b, hop) == 0
    folded = deep.launch_count - l0
    l0 = flat.launch_count
    assert flat.process_device_multi(d.data_ptr(), calls, b, hop) == 0
    unfolded = flat.launch_count - l0
    for c in range(calls):
        assert one.process_device(d.data_ptr() + 8 * c * b * hop, b, hop) == 0
    assert deep.host_feed_stats()["ring_rows"] >= calls * b > 1024
    assert flat.host_feed_stats()["ring_rows"] == 1024
    assert folded <= 8 < unfolded, (folded, unfolded)          # 2 launches + table upload vs 2 per call
    outs = [e.finish()[1] for e in (deep, flat, one)]
    for key in ("waterfall", "histogram", "spectrum"):
        assert np.array_equal(outs[0][key], outs[1][key]), key
        assert np.array_equal(outs[0][key], outs[2][key]), key
    assert deep.waterfall_position == flat.waterfall_position == (calls * b) % 1024
    # keep going on the grown ring with another geometry; still identical
    x2 = signals.noise_tones(n * 48 * 5, seed=78)
    d2 = _to_dev(torch, x2)
    for e in (deep, flat):
        assert e.process_device_multi(d2.data_ptr(), 5, 48, n) == 0
    a, bb = deep.finish()[1], flat.finish()[1]
    for key in ("waterfall", "histogram", "spectrum"):
        assert np.array_equal(a[key], bb[key]), key
    for e in (deep, flat, one):
        e.close()


def test_finish_new_rows_refreshes_only_what_changed(torch_cuda):
    """fosphor_cu_finish_new_rows copies the rows written since the last finish into the caller's
    persistent image; the image always equals a full read-back."""
    torch = torch_cuda
    n = 1024
    a = _engine()
    b = _engine()
    x = signals.noise_tones(n * (64 + 16 + 1024 + 48 + 1024 + 1024), seed=5)
    pos = 0
    rc, img, r0, nr = a.finish_new_rows()          # BOOTING: clears, everything is new
    assert (rc, r0, nr) == (1, 0, 1024)
    for size in (64, 16, 1024, 48, 2048):
        todo = size
        while todo:
            s = min(todo, 1024)
            for e in (a, b):
                assert e.process(x[pos:pos + s * n]) == 0
            pos += s * n
            todo -= s
        wf_pos0 = (a.waterfall_position - min(size, 1024)) % 1024
        rc, img, r0, nr = a.finish_new_rows()
        assert rc == 1 and nr == min(size, 1024) and r0 == (wf_pos0 if nr < 1024 else 0), (size, r0, nr)
        _, full = b.finish()
        for key in ("waterfall", "histogram", "spectrum"):
            assert np.array_equal(img[key], full[key]), (size, key)
    assert a.finish_new_rows()[0] == 0             # READY: nothing new
    a.close()
    b.close()


def test_pageable_host_feed_matches_page_locked(torch_cuda, monkeypatch):
    """The unmodified sink hands pageable memory (lib/fifo.cc:17-21).  FOSPHOR_B200_HOSTREG=0: staged by
    the copy threads in pieces; =1: the engine page-locks the caller's ring on first sight and DMAs in
    place; default (automatic): staged on first sight, registered on second sight where the process can
    read its physical page numbers (else always staged).  Same bits as a page-locked source in every
    mode; the source may be scribbled on as soon as the call returns."""
    torch = torch_cuda
    n, sizes = 1024, (1024, 1024, 256, 1024, 16, 1024, 1024, 1024)
    x = signals.noise_tones(n * sum(sizes), seed=31)
    pinned = torch.from_numpy(x.view(np.float32).copy()).pin_memory()
    outs, stats = [], {}
    for mode in ("staged", "pinned", "hostreg", "auto"):
        monkeypatch.delenv("FOSPHOR_B200_HOSTREG", raising=False)
        if mode == "staged":
            monkeypatch.setenv("FOSPHOR_B200_HOSTREG", "0")
        elif mode == "hostreg":
            monkeypatch.setenv("FOSPHOR_B200_HOSTREG", "1")
        e = _engine()
        ring = np.empty(2 * 1024 * n, np.complex64)            # a long-lived pageable "FIFO"
        pos = 0
        for i, s in enumerate(sizes):
            if mode == "pinned":
                assert e.process_host_ptr(pinned.data_ptr() + 8 * pos, s * n) == 0
            else:
                off = (i % 2) * 1024 * n
                ring[off:off + s * n] = x[pos:pos + s * n]
                assert e.process_host_ptr(ring.ctypes.data + 8 * off, s * n) == 0
                ring[off:off + s * n] = 1e9                      # recycled at once (base_sink_c_impl.cc:174)
            pos += s * n
        _, h = e.finish()
        outs.append({k: v.copy() for k, v in h.items()})
        stats[mode] = e.host_feed_stats()
        e.close()
        del ring
    for o in outs[1:]:
        for key in ("waterfall", "histogram", "spectrum"):
            assert np.array_equal(outs[0][key], o[key]), key
    calls = len(sizes)
    assert stats["staged"]["staged_calls"] == calls and stats["staged"]["direct_calls"] == 0
    assert stats["staged"]["copy_threads"] >= 1                # 8 MiB calls went through the pool
    assert stats["pinned"]["direct_calls"] == calls and stats["pinned"]["staged_calls"] == 0
    # forced: the two halves of the ring are registered on first sight (merged into one hull), every
    # later call - also the small ones - lies inside
    assert stats["hostreg"]["direct_calls"] == calls and stats["hostreg"]["staged_calls"] == 0
    assert stats["auto"]["direct_calls"] + stats["auto"]["staged_calls"] == calls
    if os.geteuid() == 0:                                      # PFNs readable: second sight was registered
        assert stats["auto"]["direct_calls"] >= 3, stats["auto"]


def test_automatic_host_registration_survives_freed_and_reallocated_buffers(torch_cuda, monkeypatch):
    """The hazard of page-locking caller memory: the caller unmaps a registered buffer, a new mapping
    lands at the same address, the stale registration would DMA the old (still pinned) pages.  The
    automatic mode compares physical page numbers before every use: results must equal the staged
    path's however often the buffer is replaced.  Buffers are real mmap()s, unmapped after each
    generation (an 8 MiB malloc may be served from the heap and keep its pages)."""
    import mmap
    n, b = 1024, 1024
    size = 8 * b * n
    monkeypatch.delenv("FOSPHOR_B200_HOSTREG", raising=False)
    auto = _engine()
    monkeypatch.setenv("FOSPHOR_B200_HOSTREG", "0")
    staged = _engine()
    addrs, reused = [], 0
    for gen in range(6):
        mm = mmap.mmap(-1, size, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        buf = np.frombuffer(mm, dtype=np.complex64)
        ptr = buf.ctypes.data
        reused += ptr in addrs
        addrs.append(ptr)
        for rep in range(3):                                   # seen, registered (if PFNs are readable), used
            buf[:] = signals.noise_tones(b * n, seed=100 + 10 * gen + rep)
            # odd generations: the staging engine meets the (possibly stale) registration of the other
            # engine FIRST - it must not trust page-locked memory that another engine of the library locked
            for e in ((auto, staged) if gen % 2 == 0 else (staged, auto)):
                assert e.process_host_ptr(ptr, b * n) == 0
        ha = {k: v.copy() for k, v in auto.finish()[1].items()}
        hs = staged.finish()[1]
        for key in ("waterfall", "histogram", "spectrum"):
            assert np.array_equal(ha[key], hs[key]), (gen, key)
        del buf
        mm.close()                                             # munmap: the registration (if any) is now stale
    st = auto.host_feed_stats()
    print("mappings that landed on an earlier address: %d of 5; automatic mode: %s" % (reused, st))
    if os.geteuid() == 0 and reused:
        # stale registrations were found and dropped: those calls were staged, not DMA'd from dead pages
        assert st["staged_calls"] >= 1 + reused, st
    auto.close()
    staged.close()


def test_export_maxhold_on_side_stream(torch_cuda):
    """fosphor_cu_export_maxhold_on: same trace as export_maxhold, without joining the engine's
    streams; the next accumulate launch waits for the read."""
    torch = torch_cuda
    n, k, b, calls, hop = 1024, 256, 1024, 8, 256
    raw = signals.noise_tones((calls * b - 1) * hop + n, seed=9)
    d = _to_dev(torch, raw)
    side = torch.cuda.Stream()
    got = []
    for use_side in (False, True):
        e = _engine(n_bins=k, wf_rows=1024)
        traces = []
        for rep in range(3):
            out = torch.empty(n, dtype=torch.float32, device="cuda")
            assert e.process_device_multi(d.data_ptr(), calls, b, hop) == 0
            if use_side:
                assert e.export_maxhold_on(out.data_ptr(), side.cuda_stream) == 0
            else:
                assert e.export_maxhold(out.data_ptr()) == 0
            traces.append(out)
        e.sync()
        side.synchronize()
        got.append([t.cpu().numpy() for t in traces])
        _, h = e.finish()
        assert np.array_equal(h["spectrum"][1, :, 1], got[-1][-1])
        e.close()
    for a, bb in zip(*got):
        assert np.array_equal(a, bb)


def test_bin_count_limits(torch_cuda):
    """K = 1024 (BASELINE configs[3]) is the largest bin count: both accumulate paths are run at it
    (fused: B = 1024; split kernels: a batch that is not a multiple of 16) and compared with the
    oracle; larger K is refused at create instead of failing at the first launch."""
    from gr_fosphor_b200.engine import Fosphor
    with pytest.raises(RuntimeError):
        Fosphor(n_bins=1025)
    with pytest.raises(RuntimeError):
        Fosphor(n_bins=2048)
    n, k = 1024, 1024
    cfg = dict(fft_len=n, n_bins=k, wf_rows=1024, batch_mult=8)
    sizes = (1024, 24, 520)
    x = signals.noise_tones(n * sum(sizes), seed=61, sigma=0.03)
    eng = Fosphor(**cfg)
    orc = oracle_lib.Oracle(**cfg)
    tw = parity.DisplayTwin(**cfg)
    pos = 0
    for s in sizes:
        p0 = eng.waterfall_position
        assert eng.process(x[pos:pos + s * n]) == 0 and orc.process(x[pos:pos + s * n]) == 0
        _, host = eng.finish()
        tw.feed(host["waterfall"], p0, s)
        pos += s * n
    orc.finish()
    tw.check(host["histogram"], host["spectrum"])
    sc, of = oracle_lib.power_range(n, 0, 10)
    parity.check_end_to_end(host, orc, np.arange(1024), sum(sizes) * n, np.float32(sc) * np.float32(k), of)
    eng.close()


def test_c_helpers_for_window_and_power_range(torch_cuda):
    """fosphor_cu_default_window / fosphor_cu_power_range (fosphor.c:108-152 for any N) against the
    oracle's restatement, bit for bit."""
    from gr_fosphor_b200 import engine
    for n in (512, 1024, 4096, 16384):
        assert np.array_equal(engine.default_window(n), oracle_lib.default_window(n))
        for ref, div in ((0, 10), (-10, 5), (10, 12)):
            assert engine.power_range(n, ref, div) == oracle_lib.power_range(n, ref, div)


def test_two_engines_on_two_devices_in_one_process(torch_cuda):
    """Several sinks per process (lib/base_sink_c_impl.cc:46,97) on several GPUs: every entry point
    sets the engine's device, kernel attributes are per device.  Interleaved calls from one thread,
    the caller's current device left alone; both must match a single-device run bit for bit."""
    torch = torch_cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n, k, b, calls = 16384, 1024, 256, 3            # 193 KB of opt-in shared memory in the FFT kernel
    x = signals.noise_tones(n * b * calls, n_fft=n, seed=13, sigma=0.02)
    torch.cuda.set_device(0)
    e0 = _engine(fft_len=n, n_bins=k, wf_rows=1024, device=0)
    e1 = _engine(fft_len=n, n_bins=k, wf_rows=1024, device=1)
    assert torch.cuda.current_device() == 0
    d0 = _to_dev(torch, x, "cuda:0")
    d1 = _to_dev(torch, x, "cuda:1")
    for c in range(calls):
        off = 8 * c * b * n
        assert e1.process_device(d1.data_ptr() + off, b) == 0
        assert e0.process_device(d0.data_ptr() + off, b) == 0
        assert torch.cuda.current_device() == 0
    h1 = {k2: v.copy() for k2, v in e1.finish()[1].items()}
    h0 = e0.finish()[1]
    for key in ("waterfall", "histogram", "spectrum"):
        assert np.array_equal(h0[key], h1[key]), key
    # host-fed on the second device while device 0 is current
    e2 = _engine(device=1)
    y = signals.noise_tones(1024 * 1024, seed=14)
    assert e2.process(y) == 0
    o = oracle_lib.Oracle()
    o.process(y)
    _, h2 = e2.finish()
    o.finish()
    parity.check_waterfall(h2["waterfall"], o.waterfall)
    for e in (e0, e1, e2):
        e.close()


def test_cfg4_channels_across_gpus_with_nccl_maxhold_reduce(torch_cuda):
    """BASELINE.json configs[3]: one independent channel per GPU and one NCCL all-reduce(MAX) of the
    max-hold trace per step; the reduced trace must be the element-wise maximum of the per-rank
    traces (tools/run_cfg4_multi.py under torchrun, as many ranks as the box has GPUs, <= 8)."""
    torch = torch_cuda
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tools", "run_cfg4_multi.py"), "4"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    assert res["n_gpus"] == world
    assert res["reduced_maxhold_is_elementwise_max"] is True


@pytest.mark.parametrize("seed", range(24))
def test_random_engine_parameters_display_stage_exact(torch_cuda, seed):
    """Seeded random walk over the parameter space the ABI accepts - FFT size, bin count (odd ones
    too), rise / decay / live constants, batch multiples of 1 / 8 / 16, call sequences that wrap the
    ring - with the strongest check there is: the engine's histogram must equal, bit for bit, what the
    oracle's display stage makes of the engine's own log-power rows (tier 1), and waterfall / flips /
    histogram / live / max-hold must agree with the oracle end to end (tier 2)."""
    from gr_fosphor_b200.engine import Fosphor
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.choice([512, 1024, 2048, 4096, 8192, 16384]))
    k = int(rng.choice([2, 3, 17, 64, 100, 128, 255, 256, 300, 512, 777, 1024]))
    mult = int(rng.choice([1, 8, 16]))
    w = int(rng.choice([256, 1024]))
    cfg = dict(fft_len=n, n_bins=k, wf_rows=w, batch_mult=mult, batch_max=256,
               t0r=float(rng.choice([4.0, 16.0, 50.0])), t0d=float(rng.choice([20.0, 256.0, 1024.0])),
               alpha=float(rng.choice([0.002, 0.01, 0.1])))
    max_units = max(1, (48 if n >= 8192 else 200) // mult)
    sizes = [int(rng.integers(1, max_units + 1)) * mult for _ in range(int(rng.integers(2, 6)))]
    x = signals.noise_tones(n * sum(sizes), n_fft=n, seed=2000 + seed, sigma=float(rng.choice([0.003, 0.02, 0.1])))
    eng = Fosphor(**cfg)
    orc = oracle_lib.Oracle(**cfg)
    tw = parity.DisplayTwin(**cfg)
    pos = 0
    for s in sizes:
        p0 = eng.waterfall_position
        assert eng.process(x[pos:pos + s * n]) == 0 and orc.process(x[pos:pos + s * n]) == 0
        _, host = eng.finish()
        tw.feed(host["waterfall"], p0, s)
        pos += s * n
    orc.finish()
    assert eng.waterfall_position == orc.waterfall_position == sum(sizes) % w
    tw.check(host["histogram"], host["spectrum"])
    sc, of = oracle_lib.power_range(n, 0, 10)
    rows = np.arange(w) if sum(sizes) >= w else np.arange(sum(sizes))
    parity.check_end_to_end(host, orc, rows, sum(sizes) * n, np.float32(sc) * np.float32(k), of, t0r=cfg["t0r"])
    eng.close()
