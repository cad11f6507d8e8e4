"""world_size-2 gloo test of the multi-GPU host logic: channel sharding and
the max-hold all-reduce (the only collective on the path, BASELINE configs[3]).
The per-channel compute is the CPU oracle here; on GPUs it is the engine."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402

N, B, CHANNELS = 512, 32, 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _channel_trace(ch):
    import oracle_lib
    import signals
    o = oracle_lib.Oracle(fft_len=N, n_bins=64)
    o.process(signals.noise_tones(N * B, n_fft=N, seed=10 + ch))
    return o.spectrum[1, :, 1].copy()


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from gr_fosphor_b200 import multi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = multi.channels_for_rank(rank, world, CHANNELS)
    local = np.full(N, -np.inf, np.float32)
    for ch in mine:
        local = np.maximum(local, _channel_trace(ch))
    t = torch.from_numpy(local.copy())
    multi.reduce_maxhold(dist, t)
    el = multi.max_over_ranks(dist, float(rank + 1))
    np.save(os.path.join(out_dir, "r%d.npy" % rank), t.numpy())
    assert el == float(world)
    dist.destroy_process_group()


def test_channel_partition():
    from gr_fosphor_b200 import multi
    parts = [multi.channels_for_rank(r, 3, 8) for r in range(3)]
    assert sorted(sum(parts, [])) == list(range(8))
    assert multi.channels_for_rank(0, 1, 8) == list(range(8))
    with pytest.raises(ValueError):
        multi.channels_for_rank(3, 3, 8)


def test_maxhold_allreduce_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    expect = np.full(N, -np.inf, np.float32)
    for ch in range(CHANNELS):
        expect = np.maximum(expect, _channel_trace(ch))
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "r%d.npy" % r))
        assert np.array_equal(got, expect)
