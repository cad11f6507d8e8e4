#!/usr/bin/env python3
"""bench.py - fosphor spectral hot path throughput on B200.

  python bench.py --gpus N --steps K --warmup W            (this repo's CUDA engine)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's own path)

Workload = BASELINE.json configs[1]: N=1024 FFT, overlap 4, continuous 100 Msps synthetic IQ
in calls of B=1024 spectra (the pre-overlapped stream fosphor_cl_process() receives; overlap_cc
upstream, lib/overlap_cc_impl.cc:64-79).

`value` (device-resident, 256 power bins as BASELINE names): one PASS is one second of that
signal - 1e8 raw samples -> 4e8 samples into the FFT = 384 calls of 1024 spectra - and one STEP
is `--passes` (default 50) passes over two alternating 3.2 GB input buffers, so that the K timed
steps cover about a second of GPU time.  The waterfall has the reference's 1024 rows
(cl.c:430-432); how many calls one launch pair folds is the engine's business (its log-power
scratch ring), not the display's.  The timed region closes after fosphor_cu_flush() has joined
the engine's accumulate stream.  A second of this work holds the board at its power cap
(clocks.reasons: sw_power_cap, SM clock ~1.55-1.65 GHz); `burst` is the same pass timed over 20 ms
on a cool chip, before the cap bites.

`e2e` (the headline against the reference arm) mirrors `--impl reference` exactly: the seven
fosphor_cl_* symbols of the drop-in library, PAGEABLE host input as the unmodified sink hands
it (lib/fifo.cc:17-21), the reference's fixed geometry (128 bins, 1024 rows), one sink frame =
8 process calls of 1024 spectra + fosphor_cl_finish (lib/base_sink_c_impl.cc:133-175) with the
4.5 MiB result read-back, H2D of every sample inside the timed region.

Multi-GPU: one engine (one channel) per GPU, weak scaling, one NCCL max all-reduce of the
max-hold trace per pass (BASELINE.json configs[3]) on a side stream.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

N_FFT, N_BINS, OVERLAP, BATCH, CALLS_PER_PASS = 1024, 256, 4, 1024, 384
WF_ROWS = 1024            # reference geometry (cl.c:430-432)
FRAME_CALLS = 8           # one sink frame (base_sink_c_impl.cc:133-146)
REF_BINS = 128            # display.cl:96: the fosphor_cl_* boundary is fixed at 128 bins

# identical in both arms (the driver compares them)
CONFIG = {
    "workload": "cfg2: N=1024 FFT, overlap=4 (pre-overlapped stream, r=1), continuous synthetic IQ in calls of "
                "B=1024 spectra",
    "fft_len": N_FFT, "overlap": OVERLAP, "batch": BATCH, "wf_rows": WF_ROWS,
    "e2e_frame": "8 fosphor_cl_process calls of 1024 spectra + fosphor_cl_finish, pageable host input, "
                 "128 bins / 1024 waterfall rows (the boundary's fixed geometry)",
}


def algorithmic_bytes_per_call(n, k, b, r):
    """SURVEY.md 8(d) / BASELINE.md 4: input + waterfall + histogram R/W + live/max R/W + window."""
    return 8.0 * n * b * r + 4.0 * n * b + 8.0 * n * k + 32.0 * n + 4.0 * n


def fft_kernel_bytes(n, spectra, r):
    """dominant kernel (fft_power): cf32 in, f32 log-power out, window once"""
    return (8.0 * r + 4.0) * n * spectra + 4.0 * n


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock, power and throttle reasons sampled through NVML every ~5 ms during the timed region."""

    def __init__(self, index):
        self.index, self.samples, self.power, self.reasons, self.max_mhz = index, [], [], set(), None
        self._stop = threading.Event()
        self.thr = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except ValueError:
                    idx = self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = {"hw_slowdown": pynvml.nvmlClocksEventReasonHwSlowdown,
                     "hw_thermal_slowdown": pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": pynvml.nvmlClocksEventReasonSwThermalSlowdown,
                     "sw_power_cap": pynvml.nvmlClocksEventReasonSwPowerCap}

            def pump():
                while not self._stop.is_set():
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                        r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        for k, bit in names.items():
                            if r & bit:
                                self.reasons.add(k)
                    except Exception:
                        pass
                    time.sleep(0.004)
            self.thr = threading.Thread(target=pump, daemon=True)
            self.thr.start()
        except Exception as exc:
            self.reasons.add("nvml unavailable: %s" % exc)

    def stop(self):
        self._stop.set()
        if self.thr:
            self.thr.join(timeout=2)
        sm = self.samples
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_mhz_min": min(sm) if sm else None,
                "sm_max_mhz": self.max_mhz, "power_w_max": max(self.power) if self.power else None,
                "power_w_median": statistics.median(self.power) if self.power else None,
                "reasons": sorted(self.reasons), "samples": len(sm)}


def synth_stream_torch(torch, n_samples, seed, device):
    """noise sigma=0.01 + 8 tones, generated on the device"""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    x = torch.randn((n_samples, 2), generator=g, device=device, dtype=torch.float32) * (0.01 / np.sqrt(2.0))
    rng = np.random.default_rng(seed)
    t = torch.arange(n_samples, device=device, dtype=torch.float64)
    for f, a, p in zip(rng.uniform(-0.5, 0.5, 8), np.exp(rng.uniform(np.log(0.02), np.log(0.6), 8)),
                       rng.uniform(0, 2 * np.pi, 8)):
        ph = (torch.remainder(t * float(f), 1.0) * (2 * np.pi) + float(p)).to(torch.float32)
        x[:, 0] += float(a) * torch.cos(ph)
        x[:, 1] += float(a) * torch.sin(ph)
        del ph
    del t
    return x


def cpu_baseline_port(seconds_budget=12.0, n_bins=N_BINS):
    """CPU oracle (f32 FFT variant, OpenMP over all host threads) on a bounded
    sample of the same workload: calls of 1024 spectra."""
    import oracle_lib
    import signals
    orc = oracle_lib.Oracle(fft_len=N_FFT, n_bins=n_bins, wf_rows=1024, fft_f32=True)
    x = signals.noise_tones(N_FFT * BATCH, seed=2).astype(np.complex64)
    orc.process(x)                       # warm-up call
    t0, calls = time.perf_counter(), 0
    while True:
        orc.process(x)
        calls += 1
        el = time.perf_counter() - t0
        if (calls >= 8 and el > seconds_budget) or calls >= 4096:
            break
    orc.finish()
    msps = calls * BATCH * N_FFT / el / 1e6
    threads = oracle_lib.lib().fosphor_oracle_threads()
    return {"value": msps, "unit": "Mcomplex-samples/s", "cores": int(threads), "kind": "port",
            "sample": "%d calls of %d spectra (N=%d, K=%d), f32 FFT oracle, %.1f s" % (calls, BATCH, N_FFT, n_bins, el)}


def bind_cpus(local, local_world):
    """Best effort: keep this rank's threads (and the first touch of its page-locked buffers) on the
    cores next to its GPU - or, when the platform reports no locality, on its own share of the
    cores - so that 8 ranks do not fight over the same ones."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        import torch
        prop = torch.cuda.get_device_properties(local)
        near, node = None, -1
        try:
            bdf = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
            base = "/sys/bus/pci/devices/" + bdf
            node = int(open(base + "/numa_node").read())
            if node >= 0:
                cl = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
                near = set()
                for part in cl.split(","):
                    lo, _, hi = part.partition("-")
                    near.update(range(int(lo), int(hi or lo) + 1))
        except Exception:
            near = None
        pool = [c for c in allowed if near is None or c in near] or allowed
        # ranks that share a pool split it
        per = max(1, len(pool) // max(1, local_world))
        mine = pool[(local * per) % len(pool):(local * per) % len(pool) + per] or pool
        if local_world > 1:
            os.sched_setaffinity(0, mine)
        return {"numa_node": node, "cpus": "%d-%d (%d)" % (mine[0], mine[-1], len(mine)) if local_world > 1
                else "all %d" % len(allowed)}
    except Exception as exc:
        return {"error": str(exc)}


def dist_setup():
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    return world, rank, local, dist


def run_b200(args):
    import torch
    from gr_fosphor_b200 import build
    from gr_fosphor_b200.dropin import FosphorCL
    from gr_fosphor_b200.engine import Fosphor

    world, rank, local, dist = dist_setup()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    cpu_bind = bind_cpus(local, local_world)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # a real (non-default) stream: the engine enqueues on it and torch.cuda.Event times it
    stream = torch.cuda.Stream(device=dev)
    comm = torch.cuda.Stream(device=dev) if dist is not None else None
    torch.cuda.set_stream(stream)
    peak, peak_src = measured_peaks()

    n, k, b, calls = N_FFT, N_BINS, BATCH, CALLS_PER_PASS
    passes = args.passes
    spectra_per_pass = calls * b
    samples_per_pass = spectra_per_pass * n
    samples_per_step = samples_per_pass * passes
    hop = n // OVERLAP

    def make_engine(mode=None, **kw):
        """mode None: the engine's default schedule (automatic: two streams when its log-power ring
        holds four chunks of >= 32 M samples); "0": every kernel on one stream (kernels
        timed alone, clean per-kernel event timing)."""
        old = os.environ.get("FOSPHOR_B200_OVERLAP")
        if mode is None:
            os.environ.pop("FOSPHOR_B200_OVERLAP", None)
        else:
            os.environ["FOSPHOR_B200_OVERLAP"] = mode
        try:
            cfg = dict(fft_len=n, n_bins=k, wf_rows=args.wf_rows, device=local, stream=stream.cuda_stream)
            cfg.update(kw)
            return Fosphor(**cfg)
        finally:
            if old is None:
                os.environ.pop("FOSPHOR_B200_OVERLAP", None)
            else:
                os.environ["FOSPHOR_B200_OVERLAP"] = old

    eng = make_engine()

    # ---- inputs: two distinct seconds of signal, raw and pre-overlapped (3.2 GB each >> L2) ----
    raw_len = (spectra_per_pass - 1) * hop + n
    raws = [synth_stream_torch(torch, raw_len, 1000 * rank + 2 + i, dev) for i in range(2)]
    # what overlap_cc would emit: [spectra][N] windows hopping N/4
    pool = [r.unfold(0, n, hop).permute(0, 2, 1).contiguous().view(-1, 2) for r in raws]
    maxhold = torch.empty(n, dtype=torch.float32, device=dev)
    torch.cuda.synchronize()

    def reduce_maxhold(e):
        """BASELINE configs[3]: reduced max-hold, on a side stream - the engine's FFT / accumulate
        pipeline is not drained for it (fosphor_cu_export_maxhold_on)"""
        e.export_maxhold_on(maxhold.data_ptr(), comm.cuda_stream)
        with torch.cuda.stream(comm):
            dist.all_reduce(maxhold, op=dist.ReduceOp.MAX)

    def timed(e, one_pass, steps, warmup, n_passes):
        """K steps of n_passes passes; CUDA events on the engine's stream, after the accumulate
        stream (and the comm stream) have been joined into it; max over ranks."""
        def step(i):
            for q in range(n_passes):
                one_pass(e, i * n_passes + q)
                if dist is not None:
                    reduce_maxhold(e)
        for i in range(warmup):
            step(i)
        e.flush()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = e.launch_count
        e0.record(stream)
        for i in range(steps):
            step(warmup + i)
        e.flush()                                   # order `stream` after the accumulate stream
        if comm is not None:
            stream.wait_stream(comm)
        e1.record(stream)
        torch.cuda.synchronize()
        launches = e.launch_count - l0
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    def pass_pre(e, i):
        e.process_device_multi(pool[i % 2].data_ptr(), calls, b, n)

    def pass_raw(e, i):
        e.process_device_multi(raws[i % 2].data_ptr(), calls, b, hop)

    # ---- the same workload in a short window on a cool chip (round 1's 20 ms region): the burst figure.
    #      A second of this work pulls the board to its power cap (~1 kW) and the SM clock down to ~1.55 GHz,
    #      which is what the sustained headline below runs at. ----
    burst_sampler = ClockSampler(local)
    if rank == 0:
        burst_sampler.start()
    ms_burst, _ = timed(eng, pass_pre, 20, 3, 1)
    burst_clocks = burst_sampler.stop() if rank == 0 else None
    burst = {"value": world * 20 * samples_per_pass / (ms_burst * 1e-3) / 1e6, "unit": "Mcomplex-samples/s",
             "ms_per_pass": ms_burst / 20, "timed_region_s": ms_burst * 1e-3,
             "step_frac": calls * algorithmic_bytes_per_call(n, k, b, 1.0) / (ms_burst / 20 * 1e-3) / 1e9 / peak,
             "clocks": burst_clocks,
             "note": "20 passes (20 ms) right after start-up, before the power cap bites: comparable with round 1's value"}
    time.sleep(0.5)

    # ---- device-resident headline ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches = timed(eng, pass_pre, args.steps, args.warmup, passes)
    clocks = sampler.stop() if rank == 0 else None
    value = world * args.steps * samples_per_step / (ms * 1e-3) / 1e6
    ring_rows = eng.host_feed_stats()["ring_rows"]

    # per-kernel durations while co-running (default schedule): event pairs around each launch
    prof_steps, prof_passes = 2, min(passes, 5)
    eng.profile(True)
    timed(eng, pass_pre, prof_steps, 0, prof_passes)
    prof_co = eng.profile_read()
    eng.profile(False)
    step_bytes = passes * calls * algorithmic_bytes_per_call(n, k, b, 1.0)

    # ---- in-engine overlap variant (raw stream, hop = N/4), same engine ----
    side_steps, side_passes = max(2, args.steps // 4), min(passes, 10)
    ms_hop, _ = timed(eng, pass_raw, side_steps, 1, side_passes)
    value_hop = world * side_steps * side_passes * samples_per_pass / (ms_hop * 1e-3) / 1e6

    # ---- the same workload with every kernel on one stream: kernels timed ALONE ----
    eng1 = make_engine("0")
    ms_one, _ = timed(eng1, pass_pre, side_steps, 1, side_passes)
    eng1.profile(True)
    timed(eng1, pass_pre, prof_steps, 0, prof_passes)
    prof = eng1.profile_read()
    eng1.profile(False)
    eng1.close()
    # ---- no folding at all: one launch pair per call of 1024 spectra (what a caller that hands the
    #      engine one call at a time gets from device-resident input) ----
    engu = make_engine(scratch_rows=-1)
    ms_unf, _ = timed(engu, pass_pre, side_steps, 1, side_passes)
    engu.close()
    per_pass = lambda t_ms: t_ms / (side_steps * side_passes)      # noqa: E731
    fft_ms = prof["fft_ms"] / max(1, prof["fft_launches"])
    count_ms = prof["count_ms"] / max(1, prof["count_launches"])
    update_ms = prof["update_ms"] / max(1, prof["update_launches"])
    prof_spectra = spectra_per_pass * prof_steps * prof_passes
    spectra_per_fft_launch = prof_spectra // max(1, prof["fft_launches"])
    fft_bytes = fft_kernel_bytes(n, spectra_per_fft_launch, 1.0)
    achieved = fft_bytes / (fft_ms * 1e-3) / 1e9
    spectra_per_fft_launch_co = prof_spectra // max(1, prof_co["fft_launches"])
    fft_ms_co = prof_co["fft_ms"] / max(1, prof_co["fft_launches"])
    acc_ms_co = (prof_co["count_ms"] / max(1, prof_co["count_launches"]) +
                 prof_co["update_ms"] / max(1, prof_co["update_launches"]))

    traffic = None
    acc_traffic_per_call = None
    tpath = os.path.join(ROOT, "profiles", "fft_power_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        # ncu capture was taken at 65536 spectra per launch: scale to this run's launch size
        traffic = tj.get("dram_bytes_per_launch") * spectra_per_fft_launch / tj.get("spectra_per_launch", 65536)
        acc_traffic_per_call = tj.get("accumulate_dram_bytes_per_call")
    pass_bytes = calls * algorithmic_bytes_per_call(n, k, b, 1.0)
    one = {"value": world * samples_per_pass / (per_pass(ms_one) * 1e-3) / 1e6,
           "unit": "Mcomplex-samples/s", "ms_per_pass": per_pass(ms_one),
           "step_frac": pass_bytes / (per_pass(ms_one) * 1e-3) / 1e9 / peak,
           "fft_share_of_kernel_time": fft_ms * prof["fft_launches"] /
           max(1e-9, prof["fft_ms"] + prof["count_ms"] + prof["update_ms"]),
           "note": "FOSPHOR_B200_OVERLAP=0: FFT and accumulate launches back to back on one stream; the per-kernel "
                   "numbers of `roofline` come from this pass (kernels timed alone)"}
    unfolded = {"value": world * samples_per_pass / (per_pass(ms_unf) * 1e-3) / 1e6, "unit": "Mcomplex-samples/s",
                "ms_per_pass": per_pass(ms_unf),
                "step_frac": pass_bytes / (per_pass(ms_unf) * 1e-3) / 1e9 / peak,
                "note": "scratch_rows=-1: the 1024-row waterfall is the ring, one FFT + one accumulate launch per "
                        "call of 1024 spectra (round-1 behaviour at the reference geometry)"}

    # ---- other BASELINE configs, device-resident, one GPU ----
    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        del pool
        torch.cuda.empty_cache()
        extras = run_extras(torch, dev, stream, local, peak, raws)
    del raws
    pool = None
    torch.cuda.empty_cache()

    # ---- e2e through the reference's own boundary, host buffers ----
    e2e = run_e2e(torch, dist, dev, local, world, args, build.LIB, FosphorCL, Fosphor, stream)

    cpu = cpu_baseline_port() if (rank == 0 and world == 1 and not args.no_cpu) else None

    if rank == 0:
        out = {
            "metric": "Mcomplex-samples/sec through FFT+histogram at N=1024",
            "value": value, "unit": "Mcomplex-samples/s",
            "spectra_per_s": value * 1e6 / n,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "realtime_factor_100Msps": 1000.0 * passes / (ms / args.steps),
            "config": CONFIG,
            "details": {"value": "device-resident, %d power bins, step = %d passes of 1 s of 100 Msps IQ "
                                 "(384 calls = 393216 spectra each)" % (k, passes),
                        "n_bins": k, "calls_per_pass": calls, "passes_per_step": passes,
                        "timed_region_s": ms * 1e-3, "wf_rows": args.wf_rows, "log_power_ring_rows": ring_rows,
                        "schedule": "engine default (automatic): two streams, chunks of ring/4 rows; "
                                    "timed region closed after fosphor_cu_flush()",
                        "l2": "each pass streams a %d MiB input (two alternating buffers), far larger than the "
                              "126 MB L2" % (samples_per_pass * 8 // 2**20),
                        "multi_gpu": ("one channel per GPU, NCCL max all-reduce of max-hold per pass on a side "
                                      "stream") if world > 1 else "single",
                        "cpu_binding": cpu_bind},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "fft_power_stream_kernel<Plan1024, TWREG>",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": traffic,
                         "timed": "kernel alone: CUDA events around each launch in the one-stream pass of the same "
                                  "workload (`one_stream`); in the default schedule it shares HBM with the "
                                  "accumulate kernel of the previous chunk, see `concurrent`",
                         "bytes_per_launch": fft_bytes, "ms_per_launch": fft_ms,
                         "spectra_per_launch": spectra_per_fft_launch,
                         "accumulate_kernel": "accumulate_fused_kernel<COLS=8, 16 counter + 8 updater warps, 256-row "
                                              "TMA boxes> (count + rise/decay + live + max-hold, one launch)",
                         "accumulate_ms_per_launch": count_ms + update_ms,
                         "accumulate_achieved_GBps": 4.0 * n * spectra_per_fft_launch / ((count_ms + update_ms) * 1e-3) / 1e9,
                         "concurrent": {"schedule": "FFT of chunk c+1 beside the accumulate kernel of chunk c (two streams)",
                                        "spectra_per_launch": spectra_per_fft_launch_co,
                                        "fft_ms_per_launch": fft_ms_co, "accumulate_ms_per_launch": acc_ms_co,
                                        "fft_achieved_GBps": fft_kernel_bytes(n, spectra_per_fft_launch_co, 1.0) / (fft_ms_co * 1e-3) / 1e9,
                                        "step_dram_GBps": None if traffic is None or acc_traffic_per_call is None else
                                        (traffic / spectra_per_fft_launch * spectra_per_pass + acc_traffic_per_call * calls) /
                                        (ms / args.steps / passes * 1e-3) / 1e9},
                         "step_algorithmic_GBps": step_bytes / (ms / args.steps * 1e-3) / 1e9,
                         "step_frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak,
                         "step_note": "the step runs for ~1 s at the board's power cap (clocks.sm_mhz, reasons: "
                                      "sw_power_cap); `burst` is the same step before the cap bites. `peak` is the "
                                      "burst copy bandwidth of MEASURED_PEAKS.json in both"},
            "e2e": e2e,
            "burst": burst,
            "one_stream": one,
            "unfolded": unfolded,
            "overlap_in_engine": {"value": value_hop, "unit": "Mcomplex-samples/s",
                                  "note": "raw stream, hop=N/4 addressing inside the FFT kernel (r=1/4), device-resident"},
        }
        if extras is not None:
            out["configs"] = extras
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def run_extras(torch, dev, stream, local, peak, raws):
    """cfg3, cfg4 (one channel), the cfg5 size sweep and the two stress inputs (constant DC = every
    hit of a column in one bin: worst-case shared-atomic contention; all zeros = the -inf path),
    device-resident on one GPU, each with its own fraction of the HBM roofline (SURVEY 8d bytes)."""
    from gr_fosphor_b200.engine import Fosphor

    def run(name, n, k, ov, b, calls, in_engine, fill=None, steps=6, **kw):
        hopv = n // ov if in_engine else n
        spectra = calls * b
        raw_len = (spectra - 1) * hopv + n
        bufs = []
        if fill is not None:
            x = torch.empty((raw_len, 2), device=dev, dtype=torch.float32)
            x[:, 0], x[:, 1] = fill
            bufs = [x, x]
        elif n == N_FFT and in_engine and raw_len <= raws[0].shape[0]:
            bufs = raws
        else:
            g = torch.Generator(device=dev)
            g.manual_seed(5)
            for _ in range(2):
                x = torch.randn((raw_len, 2), generator=g, device=dev, dtype=torch.float32) * 0.01
                x[:, 0] += 0.3 * torch.cos(torch.arange(raw_len, device=dev, dtype=torch.float32) * 0.37)
                bufs.append(x)
        eng = Fosphor(fft_len=n, n_bins=k, wf_rows=1024, device=local, stream=stream.cuda_stream, **kw)
        torch.cuda.synchronize()
        for i in range(3):
            eng.process_device_multi(bufs[i % 2].data_ptr(), calls, b, hopv)
        eng.flush()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            eng.process_device_multi(bufs[i % 2].data_ptr(), calls, b, hopv)
        eng.flush()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        eng.profile(True)
        for i in range(2):
            eng.process_device_multi(bufs[i % 2].data_ptr(), calls, b, hopv)
        prof = eng.profile_read()
        eng.profile(False)
        ring = eng.host_feed_stats()["ring_rows"]
        eng.close()
        r = 1.0 / ov if in_engine else 1.0
        gbps = calls * algorithmic_bytes_per_call(n, k, b, r) / ms / 1e6
        fft_us = prof["fft_ms"] / max(1, prof["fft_launches"]) * 1e3
        acc_us = (prof["count_ms"] / max(1, prof["count_launches"]) +
                  prof["update_ms"] / max(1, prof["update_launches"])) * 1e3
        spl = 2 * spectra / max(1, prof["fft_launches"])
        res = {"config": name, "fft_len": n, "n_bins": k, "overlap": ov, "batch": b, "calls_per_step": calls,
               "in_engine_overlap": in_engine, "r": r, "log_power_ring_rows": ring,
               "Msamples_per_s": spectra * n / ms / 1e3, "spectra_per_s": spectra / ms * 1e3,
               "ms_per_step": ms, "algorithmic_GBps": gbps, "frac": gbps / peak,
               "fft_us_per_launch": fft_us, "accumulate_us_per_launch": acc_us,
               "fft_kernel_GBps": fft_kernel_bytes(n, spl, r) / (fft_us * 1e-6) / 1e9 if fft_us else None,
               "fft_kernel_frac": fft_kernel_bytes(n, spl, r) / (fft_us * 1e-6) / 1e9 / peak if fft_us else None}
        del bufs
        torch.cuda.empty_cache()
        return res

    out = []
    # cfg3: N=4096, 512 bins, overlap 8, tau=0.95 -> t0d=20, B=256
    out.append(run("cfg3 persistence stress", 4096, 512, 8, 256, 256, True, t0d=20.0))
    # cfg4 shape on one GPU (one channel): N=16384, 1024 bins, B=1024, pre-overlapped (r=1)
    out.append(run("cfg4 one channel", 16384, 1024, 1, 1024, 16, False))
    # cfg5 sweep: K=256, overlap 4, B=1024, about 0.27 G samples per step at every size
    for n in (512, 1024, 2048, 4096, 8192, 16384):
        out.append(run("cfg5 sweep N=%d" % n, n, 256, 4, 1024, (1 << 28) // n // 1024, True))
    # cfg5 stress inputs at N=1024
    out.append(run("cfg5 stress: constant DC (all hits of a column in one bin)", 1024, 256, 4, 1024, 256, True,
                   fill=(0.25, 0.1)))
    out.append(run("cfg5 stress: all zeros (-inf path)", 1024, 256, 4, 1024, 256, True, fill=(0.0, 0.0)))
    return out


def run_e2e(torch, dist, dev, local, world, args, lib_path, FosphorCL, Fosphor, stream):
    """Host-fed figures.  `value` = the like-for-like arm (see the module docstring); the others are
    the same frames with page-locked input, with FOSPHOR_B200_HOSTREG=1, and the parameterised
    engine's own host calls (round-1 arm: K=256, 384 calls + finish from page-locked memory; raw
    stream with in-engine overlap)."""
    import signals
    n, b = N_FFT, BATCH
    frame_samples = FRAME_CALLS * b * n
    raw = signals.noise_tones((FRAME_CALLS * b - 1) * (n // OVERLAP) + n, seed=2)
    frame = signals.overlap_windows(raw, n, OVERLAP, FRAME_CALLS * b)          # pageable numpy, 64 MiB
    call_len = b * n
    os.environ["FOSPHOR_CUDA_DEV"] = str(local)

    def bracket(fn, reps, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([el], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        return el

    def dropin_frames(src_ptr, reps, env=None):
        old = {}
        for kk, vv in (env or {}).items():
            old[kk] = os.environ.get(kk)
            os.environ[kk] = vv
        try:
            eng = FosphorCL(lib_path)
        finally:
            for kk, vv in old.items():
                if vv is None:
                    os.environ.pop(kk, None)
                else:
                    os.environ[kk] = vv

        def one_frame():
            for c in range(FRAME_CALLS):
                rc = eng.process_raw(src_ptr + 8 * c * call_len, call_len)
                assert rc == 0, rc
            assert eng.finish() == 1
        el = bracket(one_frame, reps, 5)
        eng.release()
        return world * reps * frame_samples / el / 1e6, el / reps * 1e3

    frames = args.e2e_frames
    v_page, ms_frame = dropin_frames(frame.ctypes.data, frames)                       # library defaults
    v_staged, _ = dropin_frames(frame.ctypes.data, frames // 2, env={"FOSPHOR_B200_HOSTREG": "0"})
    v_reg, _ = dropin_frames(frame.ctypes.data, frames // 2, env={"FOSPHOR_B200_HOSTREG": "1"})
    pinned = torch.from_numpy(frame.view(np.float32)).pin_memory()
    v_pin, _ = dropin_frames(pinned.data_ptr(), frames // 2)
    one_thread, _ = dropin_frames(frame.ctypes.data, max(8, frames // 8),
                                  env={"FOSPHOR_B200_COPY_THREADS": "1", "FOSPHOR_B200_HOSTREG": "0"})

    # BASELINE configs[0]: one 64k-sample burst (64 spectra) in, results out - a latency, not a throughput
    try:
        burst_lat = burst_latency(lambda: FosphorCL(lib_path), frame)
    except Exception as ex:                                   # an extra; never lose the line over it
        burst_lat = "failed: %r" % (ex,)

    # the parameterised engine from page-locked memory: 48 calls of K=256 + finish, and the raw stream
    k, calls = N_BINS, 48
    eng = Fosphor(fft_len=n, n_bins=k, wf_rows=WF_ROWS, device=local, stream=stream.cuda_stream)
    hop = n // OVERLAP
    raw2 = signals.noise_tones((calls * b - 1) * hop + n, seed=3)
    h_raw = torch.from_numpy(raw2.view(np.float32)).pin_memory()
    r_wf = torch.empty((WF_ROWS, n), dtype=torch.float32).pin_memory()
    r_hist = torch.empty((k, n), dtype=torch.float32).pin_memory()
    r_spec = torch.empty((2, n, 2), dtype=torch.float32).pin_memory()

    def raw_step():
        eng.process_host_raw_ptr(h_raw.data_ptr(), calls, b, hop)
        assert eng.finish_into(r_wf.data_ptr(), r_hist.data_ptr(), r_spec.data_ptr()) == 1
    el = bracket(raw_step, 20, 3)
    v_raw = world * 20 * calls * b * n / el / 1e6
    eng.close()

    return {"value": v_page, "unit": "Mcomplex-samples/s",
            "h2d_bytes_per_step": 8 * frame_samples, "d2h_bytes_per_step": 4 * (WF_ROWS * n + REF_BINS * n + 4 * n),
            "step": "one sink frame: 8 x fosphor_cl_process(1024 spectra) + fosphor_cl_finish; %d frames timed" % frames,
            "ms_per_frame": ms_frame,
            "api": "the drop-in's fosphor_cl_* symbols (lib/fosphor/cl.h:22-32), PAGEABLE numpy input, library defaults, "
                   "128 bins, 1024 rows - the same calls, geometry and memory type as --impl reference.  Defaults: a "
                   "call range is staged by the copy threads on first sight and page-locked where it lies on second "
                   "sight, every later use checked against the physical page numbers recorded then (needs "
                   "/proc/self/pagemap with PFNs, i.e. a privileged process such as this one: root=%s; otherwise "
                   "always staged)" % (os.geteuid() == 0),
            "pcie_GBps": v_page * 8e6 / 1e9 / world,
            "cfg1_burst_latency_us": burst_lat,
            "variants": {
                "dropin_pageable_default": v_page,
                "dropin_pageable_always_staged": v_staged,
                "dropin_pageable_always_staged_one_copy_thread": one_thread,
                "dropin_pageable_hostreg_forced": v_reg,
                "dropin_page_locked_input": v_pin,
                "engine_raw_stream_in_engine_overlap": v_raw,
                "notes": "always_staged: FOSPHOR_B200_HOSTREG=0 (what an unprivileged process gets by default: copy "
                         "threads, three trips through host DRAM per byte - host-memory bound when several GPUs share "
                         "a host); hostreg_forced: FOSPHOR_B200_HOSTREG=1 (first sight, no page-number check: for "
                         "callers that vouch for their buffers, like the sink with its long-lived FIFO); page_locked: "
                         "cudaHostAlloc'ed source (the pinned FIFO of SURVEY 8f#2); raw stream: "
                         "fosphor_cu_process_host_raw, hop=N/4, each raw sample crosses PCIe once"}}


def burst_latency(make_engine, frame, reps=60):
    """BASELINE configs[0] through the boundary: fosphor_cl_process(one 64-spectrum burst of 65536 samples) +
    fosphor_cl_finish, host buffers in and out; median wall time in microseconds."""
    eng = make_engine()
    n = 64 * N_FFT
    ts = []
    for i in range(reps + 5):
        t0 = time.perf_counter()
        rc = eng.process_raw(frame.ctypes.data, n)
        assert rc == 0, rc
        assert eng.finish() == 1
        if i >= 5:
            ts.append((time.perf_counter() - t0) * 1e6)
    eng.release()
    return statistics.median(ts)


def run_reference(args):
    """The reference's own path through its own API: the unmodified cl.c +
    fft.cl + display.cl (oracle/_ref/libfosphor_ref.so) if an OpenCL device is
    reachable on this box (it then runs on the same B200 through NVIDIA's OpenCL
    - the box has no CPU OpenCL platform), else the CPU oracle port.  Host
    buffers in, host results out (fosphor_process x8 + fosphor_cl_finish per
    step).  The reference is hard-wired to 128 bins (display.cl:96)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import signals
    n, b, calls = N_FFT, BATCH, FRAME_CALLS
    raw = signals.noise_tones((calls * b - 1) * (n // OVERLAP) + n, seed=2)
    x = signals.overlap_windows(raw, n, OVERLAP, calls * b).reshape(calls, b * n)
    samples_per_step = calls * b * n
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libfosphor_ref.so")
    kind, cores, eng = None, 1, None
    if os.path.exists(ref_so):
        try:
            from gr_fosphor_b200.dropin import FosphorCL
            eng = FosphorCL(ref_so)
            kind = "reference"
        except Exception as exc:                                   # no OpenCL here
            sys.stderr.write("reference OpenCL path unavailable (%s); using the oracle port\n" % exc)
    if eng is None:
        import golden_cases
        import oracle_lib
        eng = golden_cases.OracleAdapter()
        eng.o.close()
        eng.o = oracle_lib.Oracle(fft_f32=True)
        kind, cores = "port", int(oracle_lib.lib().fosphor_oracle_threads())

    def step():
        for c in range(calls):
            assert eng.process(x[c]) == 0
        assert eng.finish() == 1

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    value = args.steps * samples_per_step / el / 1e6
    lat = None
    if kind == "reference":
        try:
            ts = []
            burst = np.ascontiguousarray(x[0][:64 * n])
            for i in range(25):
                t0 = time.perf_counter()
                assert eng.process(burst) == 0 and eng.finish() == 1
                if i >= 5:
                    ts.append((time.perf_counter() - t0) * 1e6)
            lat = statistics.median(ts)
        except Exception as ex:
            lat = "failed: %r" % (ex,)
    sample = ("%d steps of 8 calls x 1024 spectra + finish, %s" %
              (args.steps, "reference OpenCL kernels on the box's B200 via NVIDIA OpenCL (no CPU OpenCL platform exists), 128 bins"
               if kind == "reference" else "CPU oracle port, f32 FFT, all host threads, 128 bins"))
    out = {"impl": "reference", "metric": "Mcomplex-samples/sec through FFT+histogram at N=1024",
           "value": value, "unit": "Mcomplex-samples/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": CONFIG,
           "details": {"step": "one sink frame (config.e2e_frame); at N>1 rank 0 alone runs (one GPU busy)"},
           "cpu_baseline": {"value": value, "unit": "Mcomplex-samples/s", "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": "Mcomplex-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "cfg1_burst_latency_us": lat}
    if kind == "reference" and not args.no_cpu:
        # the box has no CPU OpenCL platform, so the reference's own kernels ran on the GPU; for a
        # host-cores figure next to it: the oracle port (same arithmetic, OpenMP), bounded sample
        out["cpu_port"] = cpu_baseline_port(seconds_budget=8.0, n_bins=REF_BINS)
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--passes", type=int, default=50, help="passes (1 s of signal = 384 calls each) per step")
    ap.add_argument("--e2e-frames", type=int, default=400, help="sink frames timed by the e2e arm")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg3 / cfg4 / cfg5 device-resident figures")
    ap.add_argument("--wf-rows", type=int, default=WF_ROWS, help="waterfall rows (power of two)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
