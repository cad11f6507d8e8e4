#!/usr/bin/env python3
"""bench.py - fosphor spectral hot path throughput on B200.

  python bench.py --gpus N --steps K --warmup W            (this repo's CUDA engine)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's own path)

Workload = BASELINE.json configs[1]: N=1024 FFT, 256 power bins, overlap 4,
continuous 100 Msps synthetic IQ, calls of B=1024 spectra.  One STEP is one
second of that signal: 1e8 raw samples -> 4e8 samples entering
the FFT = 384 calls of 1024 spectra (393,216 spectra, 402.7 Msamples), so
1000 / ms_per_step is the real-time headroom.  The stream is the
pre-overlapped one the reference's fosphor_cl_process() receives (overlap_cc
upstream, lib/overlap_cc_impl.cc:64-79); the in-engine-overlap variant (raw
stream, hop = N/4) is reported next to it under "overlap_in_engine".

Metric: Mcomplex-samples/s into the FFT (whole job, all GPUs).  `value` is
device-resident (inputs already in HBM, rotating over a pool larger than L2),
`e2e` goes through the C ABI with HOST buffers: H2D of every sample and the
D2H read-back of waterfall + histogram + spectrum each step inside the timed
region.  Multi-GPU: one engine (one channel) per GPU, weak scaling, one NCCL
max all-reduce of the max-hold trace per step (BASELINE.json configs[3]).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

N_FFT, N_BINS, OVERLAP, BATCH, CALLS_PER_STEP = 1024, 256, 4, 1024, 384
WF_ROWS = 262144  # device waterfall ring (1 GiB): 256 calls; the engine's two-stream schedule then works in chunks of 64 calls
REF_CALLS_PER_STEP = 8   # reference arm: one sink frame (base_sink_c_impl.cc:133-146) per step


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
    """SM clock + throttle reasons sampled through NVML every ~5 ms during the timed region."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
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
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz,
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


def cpu_baseline_port(seconds_budget=12.0):
    """CPU oracle (f32 FFT variant, OpenMP over all host threads) on a bounded
    sample of the same workload: whole steps of 8 calls x 1024 spectra."""
    import oracle_lib
    import signals
    orc = oracle_lib.Oracle(fft_len=N_FFT, n_bins=N_BINS, wf_rows=1024, fft_f32=True)
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
            "sample": "%d calls of %d spectra (N=%d, K=%d), f32 FFT oracle, %.1f s" % (calls, BATCH, N_FFT, N_BINS, el)}


def dist_setup(args):
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
    from gr_fosphor_b200.engine import Fosphor

    world, rank, local, dist = dist_setup(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # a real (non-default) stream: the engine enqueues on it and torch.cuda.Event times it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    peak, peak_src = measured_peaks()

    n, k, b, calls = N_FFT, N_BINS, BATCH, CALLS_PER_STEP
    spectra_per_step = calls * b
    samples_per_step = spectra_per_step * n
    hop = n // OVERLAP

    wf_rows = args.wf_rows

    def make_engine(rows, mode):
        """mode None: the engine's default schedule (automatic: two streams when the ring holds four
        chunks of >= 32 M samples); "0": every kernel on one stream (kernels timed
        alone, clean per-kernel event timing)."""
        old = os.environ.get("FOSPHOR_B200_OVERLAP")
        if mode is None:
            os.environ.pop("FOSPHOR_B200_OVERLAP", None)
        else:
            os.environ["FOSPHOR_B200_OVERLAP"] = mode
        try:
            return Fosphor(fft_len=n, n_bins=k, wf_rows=rows, device=local, stream=stream.cuda_stream)
        finally:
            if old is None:
                os.environ.pop("FOSPHOR_B200_OVERLAP", None)
            else:
                os.environ["FOSPHOR_B200_OVERLAP"] = old

    eng = make_engine(wf_rows, None)

    # ---- inputs: two distinct seconds of signal, raw and pre-overlapped (3.2 GB each >> L2) ----
    pool_n = raw_pool_n = 2
    raw_len = (spectra_per_step - 1) * hop + n
    raws = [synth_stream_torch(torch, raw_len, 1000 * rank + 2 + i, dev) for i in range(raw_pool_n)]
    # what overlap_cc would emit: [spectra][N] windows hopping N/4
    pool = [r.unfold(0, n, hop).permute(0, 2, 1).contiguous().view(-1, 2) for r in raws]
    maxhold = torch.empty(n, dtype=torch.float32, device=dev)
    torch.cuda.synchronize()

    def step_device(i, pre_overlapped=True):
        if pre_overlapped:
            eng.process_device_multi(pool[i % pool_n].data_ptr(), calls, b, n)
        else:
            eng.process_device_multi(raws[i % raw_pool_n].data_ptr(), calls, b, hop)
        if dist is not None:                          # BASELINE configs[3]: reduced max-hold
            eng.export_maxhold(maxhold.data_ptr())
            dist.all_reduce(maxhold, op=dist.ReduceOp.MAX)

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            fn(warmup + i)
        e1.record(stream)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident headline (clock sampling + per-kernel profiling) ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count
    ms = timed(lambda i: step_device(i, True), args.steps, args.warmup)
    launches = eng.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    launches_timed = launches * args.steps // (args.steps + args.warmup)
    value = world * args.steps * samples_per_step / (ms * 1e-3) / 1e6

    # per-kernel durations while co-running (default schedule): event pairs around each launch
    eng.profile(True)
    prof_steps = min(args.steps, 5)
    timed(lambda i: step_device(i, True), prof_steps, 0)
    prof_co = eng.profile_read()
    eng.profile(False)
    step_bytes = calls * algorithmic_bytes_per_call(n, k, b, 1.0)

    # ---- the same workload with every kernel on one stream: kernels timed ALONE ----
    eng_default = eng
    eng = make_engine(wf_rows, "0")
    ms_one = timed(lambda i: step_device(i, True), args.steps, args.warmup)
    eng.profile(True)
    timed(lambda i: step_device(i, True), prof_steps, 0)
    prof = eng.profile_read()
    eng.profile(False)
    eng.close()
    eng = eng_default
    fft_ms = prof["fft_ms"] / max(1, prof["fft_launches"])
    count_ms = prof["count_ms"] / max(1, prof["count_launches"])
    update_ms = prof["update_ms"] / max(1, prof["update_launches"])
    spectra_per_fft_launch = spectra_per_step * prof_steps // max(1, prof["fft_launches"])
    fft_bytes = fft_kernel_bytes(n, spectra_per_fft_launch, 1.0)
    achieved = fft_bytes / (fft_ms * 1e-3) / 1e9
    spectra_per_fft_launch_co = spectra_per_step * prof_steps // max(1, prof_co["fft_launches"])
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
    one = {"value": world * args.steps * samples_per_step / (ms_one * 1e-3) / 1e6,
           "unit": "Mcomplex-samples/s", "ms_per_step": ms_one / args.steps,
           "step_frac": step_bytes / (ms_one / args.steps * 1e-3) / 1e9 / peak,
           "fft_share_of_kernel_time": fft_ms * prof["fft_launches"] /
           max(1e-9, prof["fft_ms"] + prof["count_ms"] + prof["update_ms"]),
           "note": "FOSPHOR_B200_OVERLAP=0: FFT and accumulate launches back to back on one stream; the per-kernel "
                   "numbers of `roofline` come from this pass (kernels timed alone)"}

    # ---- in-engine overlap variant (raw stream, hop = N/4) ----
    ms_hop = timed(lambda i: step_device(i, False), args.steps, args.warmup)
    value_hop = world * args.steps * samples_per_step / (ms_hop * 1e-3) / 1e6

    # ---- e2e through the C ABI with host buffers ----
    h_pool = [torch.empty((samples_per_step, 2), dtype=torch.float32).pin_memory()]
    h_pool[0].copy_(pool[0])
    h_raw = [torch.empty((raw_len, 2), dtype=torch.float32).pin_memory()]
    h_raw[0].copy_(raws[0])
    torch.cuda.synchronize()
    call_len = b * n
    r_wf = torch.empty((1024, n), dtype=torch.float32).pin_memory()
    r_hist = torch.empty((k, n), dtype=torch.float32).pin_memory()
    r_spec = torch.empty((2, n, 2), dtype=torch.float32).pin_memory()

    def finish_e2e():
        rc = eng.finish_into(r_wf.data_ptr(), r_hist.data_ptr(), r_spec.data_ptr())
        assert rc == 1

    def step_e2e(i):
        base = h_pool[0].data_ptr()
        for c in range(calls):
            eng.process_host_ptr(base + 8 * c * call_len, call_len)
        if dist is not None:
            eng.export_maxhold(maxhold.data_ptr())
            dist.all_reduce(maxhold, op=dist.ReduceOp.MAX)
        finish_e2e()

    def step_e2e_raw(i):
        eng.process_host_raw_ptr(h_raw[0].data_ptr(), calls, b, hop)
        if dist is not None:
            eng.export_maxhold(maxhold.data_ptr())
            dist.all_reduce(maxhold, op=dist.ReduceOp.MAX)
        finish_e2e()

    def timed_host(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            fn(warmup + i)
        torch.cuda.synchronize()
        el = (time.perf_counter() - t0) * 1e3
        if dist is not None:
            t = torch.tensor([el], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        return el

    e2e_steps = max(2, min(args.steps, 4))
    del pool, raws
    torch.cuda.empty_cache()
    eng_dev = eng
    eng = make_engine(1024, None)    # reference-sized ring
    ms_e2e = timed_host(step_e2e, e2e_steps, 1)
    e2e = world * e2e_steps * samples_per_step / (ms_e2e * 1e-3) / 1e6
    ms_e2e_raw = timed_host(step_e2e_raw, e2e_steps, 1)
    e2e_raw = world * e2e_steps * samples_per_step / (ms_e2e_raw * 1e-3) / 1e6
    d2h = 4 * (1024 * n + k * n + 4 * n)
    eng.close()
    eng = eng_dev

    cpu = cpu_baseline_port() if (rank == 0 and world == 1 and not args.no_cpu) else None

    if rank == 0:
        out = {
            "metric": "Mcomplex-samples/sec through FFT+histogram at N=1024",
            "value": value, "unit": "Mcomplex-samples/s",
            "spectra_per_s": value * 1e6 / n,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "realtime_factor_100Msps": 1000.0 / (ms / args.steps),
            "config": {"workload": "cfg2: N=1024, 256 bins, overlap=4 (pre-overlapped stream, r=1), B=1024 "
                                   "spectra/call, step = 1 s of 100 Msps IQ = 384 calls = 393216 spectra",
                       "fft_len": n, "n_bins": k, "overlap": OVERLAP, "batch": b, "calls_per_step": calls,
                       "wf_rows": wf_rows,
                       "schedule": "engine default (automatic): two streams, chunks of wf_rows/4 rows",
                       "l2": "each step streams a %d MiB input (two alternating buffers), far larger than the 126 MB L2" % (samples_per_step * 8 // 2**20),
                       "multi_gpu": "one channel per GPU, NCCL max all-reduce of max-hold per step" if world > 1 else "single"},
            "gpu_launches": int(launches_timed),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "fft_power_stream_kernel<Plan1024, TWREG>",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": traffic,
                         "timed": "kernel alone: CUDA events around each launch in the one-stream pass of the same "
                                  "workload (`one_stream`); in the default schedule it shares HBM with the "
                                  "accumulate kernel of the previous chunk, see `concurrent`",
                         "bytes_per_launch": fft_bytes, "ms_per_launch": fft_ms,
                         "spectra_per_launch": spectra_per_fft_launch,
                         "accumulate_kernel": "accumulate_fused_kernel<COLS=8, 16 counter + 8 updater warps, 256-row TMA boxes> (count + rise/decay + live + max-hold, one launch)",
                         "accumulate_ms_per_launch": count_ms + update_ms,
                         "concurrent": {"schedule": "FFT of chunk c+1 beside the accumulate kernel of chunk c (two streams)",
                                        "spectra_per_launch": spectra_per_fft_launch_co,
                                        "fft_ms_per_launch": fft_ms_co, "accumulate_ms_per_launch": acc_ms_co,
                                        "fft_achieved_GBps": fft_kernel_bytes(n, spectra_per_fft_launch_co, 1.0) / (fft_ms_co * 1e-3) / 1e9,
                                        "step_dram_GBps": None if traffic is None or acc_traffic_per_call is None else
                                        (traffic / spectra_per_fft_launch * spectra_per_step + acc_traffic_per_call * calls) /
                                        (ms / args.steps * 1e-3) / 1e9},
                         "step_algorithmic_GBps": step_bytes / (ms / args.steps * 1e-3) / 1e9,
                         "step_frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak},
            "e2e": {"value": e2e, "unit": "Mcomplex-samples/s",
                    "h2d_bytes_per_step": 8 * samples_per_step, "d2h_bytes_per_step": d2h,
                    "api": "fosphor_cu_process_host x384 + fosphor_cu_finish (page-locked host pre-overlapped stream)"},
            "one_stream": one,
            "overlap_in_engine": {"value": value_hop, "e2e": e2e_raw, "unit": "Mcomplex-samples/s",
                                  "h2d_bytes_per_step": 8 * raw_len,
                                  "note": "raw stream, hop=N/4 addressing inside the FFT kernel (r=1/4)"},
        }
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


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
    n, b, calls = N_FFT, BATCH, REF_CALLS_PER_STEP
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
    sample = ("%d steps of 8 calls x 1024 spectra, %s" %
              (args.steps, "reference OpenCL kernels on the box's B200 via NVIDIA OpenCL (no CPU OpenCL platform exists), 128 bins"
               if kind == "reference" else "CPU oracle port, f32 FFT, all host threads, 128 bins"))
    out = {"impl": "reference", "metric": "Mcomplex-samples/sec through FFT+histogram at N=1024",
           "value": value, "unit": "Mcomplex-samples/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "cfg2: N=1024, overlap=4 (pre-overlapped stream), B=1024 spectra/call, "
                                  "step = 8 calls + finish; reference fixed at 128 bins"},
           "cpu_baseline": {"value": value, "unit": "Mcomplex-samples/s", "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": "Mcomplex-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if kind == "reference" and not args.no_cpu:
        # the box has no CPU OpenCL platform, so the reference's own kernels ran on the GPU; for a
        # host-cores figure next to it: the oracle port (same arithmetic, OpenMP), bounded sample
        out["cpu_port"] = cpu_baseline_port(seconds_budget=8.0)
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--wf-rows", type=int, default=WF_ROWS, help="device waterfall ring rows (power of two)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
