#!/usr/bin/env python3
"""Hot SASS lines of one kernel in an .ncu-rep (source page): stall samples with their reasons.
usage: ncu_hot.py REPORT.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    # one segment per kernel: a "Kernel Name" row, a header row, then the SASS lines
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    for a, b in zip(starts[:-1], starts[1:]):
        report(rows[a:b], top)


def report(rows, top):
    hdr = rows[1]
    print(rows[0][1][:120])
    isrc, ie, iss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for n, r in enumerate(rows[2:]):
        if len(r) <= ie:
            continue
        reasons = sorted(((int(r[i] or 0), name) for i, name in stall), reverse=True)[:3]
        data.append((n, r[isrc], int(r[ie] or 0), int(r[iss] or 0), reasons))
    tot_s = sum(d[3] for d in data)
    tot_e = sum(d[2] for d in data)
    print("lines %d  samples %d  warp instructions %d" % (len(data), tot_s, tot_e))
    agg = {}
    for d in data:
        for c, name in d[4]:
            agg[name] = agg.get(name, 0) + c
    print("stall mix (top-3 per line summed):", sorted(((v, k) for k, v in agg.items()), reverse=True)[:8])
    for d in sorted(data, key=lambda d: -d[3])[:top]:
        print("%5d %-70s exec %9d  samples %6d %5.1f%%  %s" % (d[0], d[1][:70], d[2], d[3], 100.0 * d[3] / max(1, tot_s),
              " ".join("%s:%d" % (n, c) for c, n in d[4] if c)))
    print()


if __name__ == "__main__":
    main()
