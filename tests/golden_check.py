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


def check_case(name, engine):
    """Replays the case on `engine` and checks every recorded step against the fixture."""
    steps = golden_cases.cases()[name]
    z, meta = load(name)
    rec = golden_cases.replay(engine, steps)
    assert len(rec) == len(meta)
    spectra_so_far = 0
    it = iter(steps)
    for r, m in zip(rec, meta):
        st = next(it)
        while st[0] not in ("process", "finish"):
            st = next(it)
        assert r["op"] == m["op"] == st[0]
        assert r["rc"] == m["rc"], (name, m)
        if r["op"] == "process":
            if r["rc"] == 0:
                spectra_so_far += st[1].size // 1024
            continue
        assert r["wf_pos"] == m["wf_pos"]
        key = m["key"]
        rows = z[key + "_wf_rows"]
        parity.check_waterfall(r["waterfall"][rows], z[key + "_waterfall"])
        parity.check_histogram(r["histogram"], z[key + "_histogram"], hits_in_play=max(1, spectra_so_far) * 1024)
        parity.check_spectrum(r["spectrum"], z[key + "_spectrum"], wf_ref=z[key + "_waterfall"])
