"""Compare a replay of tests/golden_cases.py with the committed fixtures that
were produced by the REAL reference (tests/golden/make_golden.py)."""
import json
import os

import numpy as np

import golden_cases
import parity

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


# columns that hold only rounding noise and may be left out of the live / max-hold comparison:
# the analytic tones under a rectangular window leave every off-peak bin at the f32 noise floor
MAX_SKIPPED = {"kat_tones": 1024}


def check_case(name, engine):
    """Replays the case on `engine` and checks every recorded step against the fixture.
    Returns a summary (observed bin flips, out-of-tolerance histogram cells, skipped columns)."""
    import oracle_lib
    steps = golden_cases.cases()[name]
    z, meta = load(name)
    rec = golden_cases.replay(engine, steps)
    assert len(rec) == len(meta)
    spectra_so_far = 0
    calls_ok = 0
    hscale, hofs = oracle_lib.power_range(1024, 0, 10)          # fosphor_init default (fosphor.c:66)
    it = iter(steps)
    summary = {"flips": 0, "bad_cells": 0, "hits": 0}
    for r, m in zip(rec, meta):
        st = next(it)
        while st[0] not in ("process", "finish"):
            if st[0] == "range":
                hscale, hofs = oracle_lib.power_range(1024, st[1], st[2])
            st = next(it)
        assert r["op"] == m["op"] == st[0]
        assert r["rc"] == m["rc"], (name, m)
        if r["op"] == "process":
            if r["rc"] == 0:
                spectra_so_far += st[1].size // 1024
                calls_ok += 1
            continue
        assert r["wf_pos"] == m["wf_pos"]
        key = m["key"]
        rows = z[key + "_wf_rows"]
        parity.check_waterfall(r["waterfall"][rows], z[key + "_waterfall"])
        # bin flips between the two waterfalls on the rows the fixture holds (scale * 128: cl.c:1087)
        flips = parity.count_flips(r["waterfall"][rows], z[key + "_waterfall"], hscale * 128, hofs, 128)
        hits = max(1, spectra_so_far) * 1024
        bad, _ = parity.check_histogram(r["histogram"], z[key + "_histogram"], hits_in_play=hits, flips=flips,
                                        visible_hits=max(1, len(rows)) * 1024, single_call=calls_ok == 1)
        parity.check_spectrum(r["spectrum"], z[key + "_spectrum"], wf_ref=z[key + "_waterfall"],
                              max_skipped=MAX_SKIPPED.get(name, 0))
        summary = {"flips": flips, "bad_cells": bad, "hits": hits}
    return summary
