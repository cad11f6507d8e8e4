#!/usr/bin/env python3
"""A/B timing of the accumulate stage variants (env knobs of engine.cu) on one GPU,
cfg2 shape unless told otherwise.  Prints one JSON line per variant."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)
import perf_configs  # noqa: E402


def main():
    import torch
    shapes = {"cfg2": ("cfg2", 1024, 256, 4, 1024, 64, 32768, False, {}),
              "cfg3": ("cfg3", 4096, 512, 8, 256, 256, 32768, True, {"t0d": 20.0}),
              "cfg3r8": ("cfg3r8", 4096, 512, 8, 256, 32, 8192, True, {"t0d": 20.0}),
              "cfg4": ("cfg4", 16384, 1024, 1, 1024, 32, 16384, False, {}),
              "cfg4r4": ("cfg4r4", 16384, 1024, 1, 1024, 4, 4096, False, {}),
              "n2048": ("n2048", 2048, 256, 4, 1024, 32, 16384, True, {}),
              "n4096": ("n4096", 4096, 256, 4, 1024, 32, 16384, True, {}),
              "n8192": ("n8192", 8192, 256, 4, 1024, 32, 16384, True, {}),
              "n16384": ("n16384", 16384, 256, 4, 1024, 32, 16384, True, {}),
              "cfg2r64": ("cfg2r64", 1024, 256, 4, 1024, 128, 65536, False, {}),
              "cfg2r128": ("cfg2r128", 1024, 256, 4, 1024, 256, 131072, False, {}),
              "cfg2r256": ("cfg2r256", 1024, 256, 4, 1024, 512, 262144, False, {}),
              "n4096r64": ("n4096r64", 4096, 256, 4, 1024, 128, 65536, True, {}),
              "n2048r128": ("n2048r128", 2048, 256, 4, 1024, 256, 131072, True, {}),
              "n16384r16": ("n16384r16", 16384, 256, 4, 1024, 32, 16384, True, {}),
              "cfg3r256": ("cfg3r256", 4096, 512, 8, 256, 512, 65536, True, {"t0d": 20.0}),
              "cfg4r64": ("cfg4r64", 16384, 1024, 1, 1024, 64, 65536, False, {}),
              "n512r512": ("n512r512", 512, 256, 4, 1024, 1024, 524288, True, {}),
              "cfg2r512": ("cfg2r512", 1024, 256, 4, 1024, 1024, 524288, False, {}),
              "n512": ("n512", 512, 256, 4, 1024, 128, 65536, True, {})}
    which = sys.argv[1:] or ["cfg2"]
    variants = [
        {"ACC": "1", "ACC_REP": "1"},
    ]
    # var=KEY:VAL,KEY:VAL ... : explicit variants (FOSPHOR_B200_<KEY> = VAL), "var=" is the default build
    explicit = [w for w in which if w.startswith("var=")]
    if explicit:
        which = [w for w in which if not w.startswith("var=")]
        variants = [dict(kv.split(":") for kv in w[4:].split(",") if kv) for w in explicit]
    if "chunks" in which:
        which.remove("chunks")
        variants = [{"ACC": a, "CHUNK_CALLS": c} for a in ("0", "1") for c in ("0", "16", "8", "4")]
    modes = ("0",)
    if "twostream" in which:
        which.remove("twostream")
        modes = ("1",)
        variants = [{"ACC": "1", "OVERLAP_CHUNK": c, "ACC_SLIM": s} for s in ("1", "0") for c in ("4", "8", "16", "32")]
        variants.append({"ACC": "1", "OVERLAP_CHUNK": "8", "ACC_SLIM": "1", "FFT_CTAS": "3"})
    for w in which:
        name, n, k, ov, b, calls, rows, ieo, kw = shapes[w]
        for var in variants:
            for key, v in var.items():
                os.environ["FOSPHOR_B200_" + key] = v
            for mode in ((var["OVERLAP"],) if "OVERLAP" in var else modes):
                os.environ["FOSPHOR_B200_OVERLAP"] = mode
                r = perf_configs.run_one(torch, name, n, k, ov, b, calls, rows, ieo, **kw)
                print(json.dumps({"shape": w, "variant": var, "two_stream": mode == "1",
                                  "Msps": round(r["Msamples_per_s"]), "ms_per_step": round(r["ms_per_step"], 4),
                                  "fft_us": round(r["fft_us_per_launch"], 1),
                                  "acc_us": round(r["count_us_per_launch"], 1),
                                  "upd_us": round(r["update_us_per_launch"], 1)}), flush=True)
            for key in var:
                os.environ.pop("FOSPHOR_B200_" + key)


if __name__ == "__main__":
    main()
