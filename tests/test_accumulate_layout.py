"""Host-side restatement of the fused accumulate kernel's work assignment
(gr-fosphor_b200/csrc/accumulate.cuh: acc_rows_per_vwarp, the (call, virtual warp) pair loop of the
counter warps, the TMA box rule of engine.cu:launch_accumulate_fused).  Checks, for many batch
sizes, that every row of every call is counted exactly once, that every chunk a warp waits for
lies inside one box, and that each box is released by exactly as many virtual warps as its
`empty` barrier expects."""
import pytest

ACC_VW = 16


def rows_per_vwarp(batch):
    units = (batch + 15) // 16
    rv = 16 * ((units + ACC_VW - 1) // ACC_VW)
    return max(rv, 64)


def box_rows(batch, wf_pos, wf_rows, box_max=256):
    """engine.cu: largest box that tiles the batch and the runs and never straddles the ring end"""
    rv = rows_per_vwarp(batch)
    b = box_max
    while b >= 16:
        if batch % b == 0 and wf_pos % b == 0 and wf_rows % b == 0 and (b % rv == 0 or rv % b == 0):
            return b
        b >>= 2
    return 0


@pytest.mark.parametrize("fw", [8, 16])
@pytest.mark.parametrize("gc", [1, 2, 4])
def test_every_row_counted_once(fw, gc):
    for batch in list(range(16, 1200, 16)) + [2048, 4096, 8192, 32768]:
        rv = rows_per_vwarp(batch)
        nv = (batch + rv - 1) // rv
        assert 1 <= nv <= ACC_VW
        n_calls = 2 * gc + 1                      # a short last group
        seen = {}
        n_groups = (n_calls + gc - 1) // gc
        for grp in range(n_groups):
            ncg = min(gc, n_calls - grp * gc)
            for warp in range(fw):
                p = warp
                while p < ncg * nv:
                    ci, v = divmod(p, nv)
                    call = grp * gc + ci
                    lo = v * rv
                    rows = min(batch - lo, rv)
                    assert rows >= 1
                    for row in range(lo, lo + rows):
                        assert (call, row) not in seen, (batch, call, row)
                        seen[(call, row)] = warp
                    p += fw
        assert len(seen) == n_calls * batch, batch


def test_boxes_and_sharers():
    wf_rows = 1 << 16
    for batch in list(range(16, 1200, 16)) + [2048, 4096, 32768]:
        for wf_pos in (0, 16, 64, 256, 4096 + 48):
            boxr = box_rows(batch, wf_pos, wf_rows)
            if boxr == 0:
                continue
            rv = rows_per_vwarp(batch)
            nv = (batch + rv - 1) // rv
            ch = min(rv, boxr)                     # rows per chunk a warp waits for
            sharers = boxr // rv if boxr > rv else 1
            users = {}
            for call in range(3):
                for v in range(nv):
                    lo = v * rv
                    rows = min(batch - lo, rv)
                    assert rows % ch == 0, (batch, boxr)          # no partial chunk ever reads past its run
                    for c0 in range(0, rows, ch):
                        g = call * batch + lo + c0
                        assert g // boxr == (g + ch - 1) // boxr, (batch, boxr)   # chunk inside one box
                        users.setdefault(g // boxr, set()).add((call, v))
            assert len(users) == 3 * batch // boxr
            for box, us in users.items():
                assert len(us) == sharers, (batch, boxr, box)
