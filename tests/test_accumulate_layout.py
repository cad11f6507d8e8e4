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


def ring_safe(depth, boxes_per_group, boxes_total):
    """accumulate.cuh: acc_ring_safe"""
    return boxes_total <= depth or depth % boxes_per_group == 0 or depth >= 2 * boxes_per_group


def _waits(batch, boxr, fw, gc, n_calls):
    """program of every counter warp: [(group, box index)] in order"""
    rv = rows_per_vwarp(batch)
    nv = (batch + rv - 1) // rv
    ch = min(rv, boxr)
    progs = {w: [] for w in range(fw)}
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
                for c0 in range(0, rows, ch):
                    progs[warp].append((grp, (call * batch + lo + c0) // boxr))
                p += fw
    return progs


def test_stage_ring_parity_waits_are_sound():
    """A parity wait for box n is sound only if box n - depth has landed before the wait starts in
    EVERY schedule: the same warp consumed it earlier, or it belongs to group g-2 or older (all
    counters finished that group before any warp may enter group g).  Whenever the host rule says
    "safe", that must hold for every wait of every warp; the shapes that hung (B = 32768 on a ring of
    32 boxes, B = 8192 on a ring of 16) must be rejected."""
    checked = rejected = 0
    for fw in (8, 16):
        for gc in (1, 2, 4):
            for batch in [16, 64, 256, 528, 1024, 1040, 2048, 2112, 4096, 8192, 32768]:
                boxr = box_rows(batch, 0, 1 << 20)
                if boxr == 0:
                    continue
                bpc = batch // boxr
                for n_calls in (1, 3, 9):
                    total = n_calls * bpc
                    progs = _waits(batch, boxr, fw, gc, n_calls)
                    for dlog in range(0, 10):
                        depth = 1 << dlog
                        if depth > 8192 // boxr:
                            break
                        if not ring_safe(depth, gc * bpc, total):
                            rejected += 1
                            continue
                        checked += 1
                        for warp, prog in progs.items():
                            mine = set()
                            for grp, n in prog:
                                old = n - depth
                                if old >= 0:
                                    old_grp = (old // bpc) // gc
                                    assert old in mine or old_grp <= grp - 2, (batch, boxr, fw, gc, n_calls, depth, warp, n)
                                mine.add(n)
    assert checked > 100 and rejected > 10
    assert not ring_safe(32, 128, 128)      # B = 32768, 256-row boxes, ring of 32
    assert not ring_safe(16, 32, 32)        # B = 8192, ring of 16
    assert ring_safe(16, 4, 256)            # bench: B = 1024
    assert ring_safe(4, 4, 128)             # cfg3: four calls of one box per group
