#!/usr/bin/env python3
"""Run a few device-resident steps of one configuration (for ncu captures).
usage: run_config.py N K B CALLS WF_ROWS HOP [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from gr_fosphor_b200.engine import Fosphor
    n, k, b, calls, wf, hop = (int(v) for v in sys.argv[1:7])
    steps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
    dev = torch.device("cuda", 0)
    eng = Fosphor(fft_len=n, n_bins=k, wf_rows=wf)
    raw_len = (calls * b - 1) * hop + n
    x = torch.randn((raw_len, 2), device=dev, dtype=torch.float32) * 0.01
    x[:, 0] += 0.3 * torch.cos(torch.arange(raw_len, device=dev, dtype=torch.float32) * 0.37)
    torch.cuda.synchronize()
    for _ in range(steps):
        eng.process_device_multi(x.data_ptr(), calls, b, hop)
    eng.sync()
    eng.close()


if __name__ == "__main__":
    main()
