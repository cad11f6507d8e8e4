#!/usr/bin/env python3
"""Device-resident throughput of the other BASELINE.json configs (cfg3, cfg4 shape,
cfg5 sweep) on one GPU: Msamples/s into the FFT, spectra/s and algorithmic GB/s
(SURVEY.md 8d bytes) with per-kernel CUDA-event times.  Not the bench headline
(that is cfg2, bench.py); feeds the table in BASELINE.md / profiles/.

    python tools/perf_configs.py > profiles/r1_configs.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def alg_bytes_per_call(n, k, b, r):
    return 8.0 * n * b * r + 4.0 * n * b + 8.0 * n * k + 32.0 * n + 4.0 * n


def run(torch, name, n, k, overlap, b, calls, wf_rows, in_engine_overlap, **kw):
    """the engine's default schedule (automatic two-stream for large rings) next to one stream"""
    os.environ.pop("FOSPHOR_B200_OVERLAP", None)
    best = run_one(torch, name, n, k, overlap, b, calls, wf_rows, in_engine_overlap, **kw)
    best["schedule"] = "default (automatic)"
    os.environ["FOSPHOR_B200_OVERLAP"] = "0"
    other = run_one(torch, name, n, k, overlap, b, calls, wf_rows, in_engine_overlap, **kw)
    os.environ.pop("FOSPHOR_B200_OVERLAP")
    best["one_stream_Msamples_per_s"] = other["Msamples_per_s"]
    for key in ("fft_us_per_launch", "count_us_per_launch", "update_us_per_launch", "fft_launches_per_step"):
        best["one_stream_" + key] = other[key]
    return best


def run_one(torch, name, n, k, overlap, b, calls, wf_rows, in_engine_overlap, **kw):
    from gr_fosphor_b200.engine import Fosphor
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng = Fosphor(fft_len=n, n_bins=k, wf_rows=wf_rows, stream=stream.cuda_stream, **kw)
    hop = n // overlap if in_engine_overlap else n
    spectra = calls * b
    raw_len = (spectra - 1) * hop + n
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    bufs = [torch.randn((raw_len, 2), generator=g, device=dev, dtype=torch.float32) * 0.01 for _ in range(2)]
    for x in bufs:
        x[:, 0] += 0.3 * torch.cos(torch.arange(raw_len, device=dev, dtype=torch.float32) * 0.37)
    torch.cuda.synchronize()

    def step(i):
        eng.process_device_multi(bufs[i % 2].data_ptr(), calls, b, hop)

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    steps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        step(i)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    eng.profile(True)
    for i in range(3):
        step(i)
    prof = eng.profile_read()
    eng.profile(False)
    samples = spectra * n
    r = (1.0 / overlap) if in_engine_overlap else 1.0
    out = {"config": name, "fft_len": n, "n_bins": k, "overlap": overlap, "batch": b, "calls_per_step": calls,
           "wf_rows": wf_rows, "in_engine_overlap": in_engine_overlap,
           "Msamples_per_s": samples / ms / 1e3, "spectra_per_s": spectra / ms * 1e3,
           "algorithmic_GBps": calls * alg_bytes_per_call(n, k, b, r) / ms / 1e6,
           "ms_per_step": ms,
           "fft_us_per_launch": prof["fft_ms"] / max(1, prof["fft_launches"]) * 1e3,
           "count_us_per_launch": prof["count_ms"] / max(1, prof["count_launches"]) * 1e3,
           "update_us_per_launch": prof["update_ms"] / max(1, prof["update_launches"]) * 1e3,
           "fft_launches_per_step": prof["fft_launches"] / 3}
    eng.close()
    del bufs
    torch.cuda.empty_cache()
    return out


def main():
    import torch
    peak = 6581.2
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    res = []
    # cfg2 both ways for reference
    res.append(run(torch, "cfg2 pre-overlapped", 1024, 256, 4, 1024, 512, 262144, False))
    res.append(run(torch, "cfg2 in-engine overlap", 1024, 256, 4, 1024, 512, 262144, True))
    # cfg3: N=4096, 512 bins, overlap 8, tau=0.95 -> t0d=20, B=256
    res.append(run(torch, "cfg3 persistence stress", 4096, 512, 8, 256, 512, 65536, True, t0d=20.0))
    # cfg4 shape on one GPU (one channel): N=16384, 1024 bins, B=1024
    res.append(run(torch, "cfg4 one channel", 16384, 1024, 1, 1024, 32, 16384, False))
    # cfg5 sweep: K=256, overlap 4, B=1024
    for n in (512, 1024, 2048, 4096, 8192, 16384):
        rows = (1 << 28) // n                          # a 1 GiB waterfall ring at every size
        res.append(run(torch, "cfg5 sweep N=%d" % n, n, 256, 4, 1024, (rows // 1024) * 2, rows, True))
    for r in res:
        r["frac_of_hbm_peak"] = r["algorithmic_GBps"] / peak
    print(json.dumps({"hbm_peak_GBps": peak, "results": res}, indent=1))


if __name__ == "__main__":
    main()
