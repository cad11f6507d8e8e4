#!/usr/bin/env python3
"""A/B of engine tuning knobs (FOSPHOR_B200_<KEY>) on a device-resident workload, default = cfg2
(N=1024, K=256, B=1024, 384 calls per pass).  Usage:
    python tools/ab_env.py [shape=N,K,B,CALLS,OVERLAP_IN_ENGINE] var=KEY:VAL,KEY:VAL var= ...
Prints one JSON line per variant: Msamples/s (CUDA events, flush before the closing event),
per-kernel microseconds, and whether histogram / waterfall are bit-identical to the first variant."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    from gr_fosphor_b200.engine import Fosphor
    shape = (1024, 256, 1024, 384, 0)
    variants = []
    passes = 12
    for a in sys.argv[1:]:
        if a.startswith("passes="):
            passes = int(a[7:])
        if a.startswith("shape="):
            shape = tuple(int(v) for v in a[6:].split(","))
        elif a.startswith("var="):
            variants.append(dict(kv.split(":") for kv in a[4:].split(",") if kv))
    variants = variants or [{}]
    n, k, b, calls, ieo = shape
    hop = n // 4 if ieo else n
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    raw_len = (calls * b - 1) * hop + n
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    bufs = []
    for _ in range(2):
        x = torch.randn((raw_len, 2), generator=g, device=dev, dtype=torch.float32) * 0.01
        x[:, 0] += 0.3 * torch.cos(torch.arange(raw_len, device=dev, dtype=torch.float32) * 0.37)
        bufs.append(x)
    first = None
    for var in variants:
        env = {key: v for key, v in var.items() if key != "SCRATCH"}      # SCRATCH: the scratch_rows parameter
        for key, v in env.items():
            os.environ["FOSPHOR_B200_" + key] = v
        eng = Fosphor(fft_len=n, n_bins=k, wf_rows=1024, stream=stream.cuda_stream,
                      scratch_rows=int(var.get("SCRATCH", 0)))
        for key in env:
            os.environ.pop("FOSPHOR_B200_" + key)
        for i in range(4):
            eng.process_device_multi(bufs[i % 2].data_ptr(), calls, b, hop)
        eng.flush()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(passes):
            eng.process_device_multi(bufs[i % 2].data_ptr(), calls, b, hop)
        eng.flush()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / passes
        eng.profile(True)
        for i in range(2):
            eng.process_device_multi(bufs[i % 2].data_ptr(), calls, b, hop)
        prof = eng.profile_read()
        eng.profile(False)
        _, host = eng.finish()
        sig = {key: host[key].copy() for key in ("waterfall", "histogram", "spectrum")}
        if first is None:
            first = sig
        same = {key: bool(np.array_equal(first[key], sig[key])) for key in sig}
        dspec = float(np.nanmax(np.abs(first["spectrum"] - sig["spectrum"])))
        print(json.dumps({"variant": var, "Msps": round(calls * b * n / ms / 1e3), "ms_per_pass": round(ms, 4),
                          "fft_us": round(prof["fft_ms"] / max(1, prof["fft_launches"]) * 1e3, 1),
                          "fft_launches": prof["fft_launches"] // 2,
                          "acc_us": round((prof["count_ms"] / max(1, prof["count_launches"]) +
                                           prof["update_ms"] / max(1, prof["update_launches"])) * 1e3, 1),
                          "two_stream_chunks": eng.two_stream_chunks, "same": same, "max_dspec": dspec}), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
