#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the REAL reference.

Run on a box with an OpenCL device (the B200 box: NVIDIA OpenCL, found through
dlopen by oracle/ref_build/ref_shim.c), after `make -C oracle/ref_build` in the
build container:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'

It replays tests/golden_cases.py through oracle/_ref/libfosphor_ref.so (the
reference's unmodified cl.c + fft.cl + display.cl) and stores every finish()
snapshot; waterfalls are stored as the rows written so far (or every 8th row
once the ring is full) to keep the fixtures small.  Also prints the oracle's
deviation from the reference for a first look.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import golden_cases  # noqa: E402
from gr_fosphor_b200.dropin import FosphorCL  # noqa: E402

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libfosphor_ref.so")


def wf_rows_to_keep(total_rows):
    if total_rows <= 128:
        return np.arange(total_rows)
    return np.arange(0, 1024, 32)


def pack(records, steps):
    out, meta, rows_written = {}, [], 0
    it = iter(records)
    for st in steps:
        if st[0] == "process":
            r = next(it)
            meta.append({"op": "process", "rc": r["rc"]})
            if r["rc"] == 0:
                rows_written += st[1].size // 1024
        elif st[0] == "finish":
            r = next(it)
            i = len(meta)
            rows = wf_rows_to_keep(min(rows_written, 1024) if rows_written else 1024)
            if rows_written == 0:
                rows = np.arange(0, 1024, 32)
            meta.append({"op": "finish", "rc": r["rc"], "wf_pos": r["wf_pos"], "key": "s%d" % i})
            out["s%d_wf_rows" % i] = rows.astype(np.int32)
            out["s%d_waterfall" % i] = r["waterfall"][rows]
            out["s%d_histogram" % i] = r["histogram"]
            out["s%d_spectrum" % i] = r["spectrum"]
    return out, meta


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    summary = {}
    for name, steps in golden_cases.cases().items():
        ref = FosphorCL(REF_SO)
        rec_ref = golden_cases.replay(ref, steps)
        ref.release()
        arrays, meta = pack(rec_ref, steps)
        arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **arrays)

        # first look: oracle vs reference
        rec_or = golden_cases.replay(golden_cases.OracleAdapter(), steps)
        worst = {}
        for a, b in zip(rec_ref, rec_or):
            assert a["op"] == b["op"]
            if a["rc"] != b["rc"]:
                worst["rc_mismatch"] = worst.get("rc_mismatch", 0) + 1
            if a["op"] != "finish":
                continue
            for k in ("waterfall", "histogram", "spectrum"):
                x, y = a[k], b[k]
                same_nonfinite = (~np.isfinite(x)) & (~np.isfinite(y)) & ((x == y) | (np.isnan(x) & np.isnan(y)))
                d = np.where(same_nonfinite, 0.0, np.abs(x.astype(np.float64) - y.astype(np.float64)))
                d = np.nan_to_num(d, nan=np.inf)
                worst[k] = max(worst.get(k, 0.0), float(d.max()))
                if k == "histogram":
                    worst["hist_cells_gt_2e-3"] = worst.get("hist_cells_gt_2e-3", 0) + int((d > 2e-3).sum())
        summary[name] = worst
        print(name, json.dumps(worst))
    with open(os.path.join(outdir, "summary_oracle_vs_reference.json"), "w") as f:
        json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main()
