#!/usr/bin/env python3
"""BASELINE.json configs[3]: N=16384, 1024 bins, B=1024, one independent channel per GPU,
NCCL all-reduce(MAX) of the max-hold trace after every step.  Launch under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29512 tools/run_cfg4_multi.py [steps]

Rank 0 prints one JSON line: whole-job Msamples/s (device-resident input, CUDA events, max over
ranks) and a check that the reduced trace is the element-wise max of the per-rank traces."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    from gr_fosphor_b200.engine import Fosphor
    from gr_fosphor_b200 import multi

    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, k, b, calls, rows = 16384, 1024, 1024, 16, 1024
    stream = torch.cuda.Stream(device=dev)
    comm = torch.cuda.Stream(device=dev)           # export + all-reduce: never joins the engine's pipeline
    torch.cuda.set_stream(stream)
    eng = Fosphor(fft_len=n, n_bins=k, wf_rows=rows, device=local, stream=stream.cuda_stream)
    g = torch.Generator(device=dev)
    g.manual_seed(10 + rank)                       # channel seeds 10..17
    samples = calls * b * n
    bufs = []
    for _ in range(2):
        x = torch.randn((samples, 2), generator=g, device=dev, dtype=torch.float32) * 0.01
        t = torch.arange(samples, device=dev, dtype=torch.float32)
        x[:, 0] += 0.3 * torch.cos(t * (0.11 + 0.07 * rank))      # a distinct tone per channel
        bufs.append(x)
        del t
    trace = torch.empty(n, dtype=torch.float32, device=dev)
    own = torch.empty(n, dtype=torch.float32, device=dev)

    def step(i):
        eng.process_device_multi(bufs[i % 2].data_ptr(), calls, b, n)
        eng.export_maxhold_on(trace.data_ptr(), comm.cuda_stream)
        with torch.cuda.stream(comm):
            own.copy_(trace)
            multi.reduce_maxhold(dist if world > 1 else None, trace)

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        step(i)
    eng.flush()
    stream.wait_stream(comm)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = multi.max_over_ranks(dist if world > 1 else None, e0.elapsed_time(e1), device=dev)
    ok = True
    if world > 1:
        gathered = [torch.empty_like(own) for _ in range(world)]
        dist.all_gather(gathered, own)
        ok = bool(torch.equal(torch.stack(gathered).max(dim=0).values, trace))
    if rank == 0:
        print(json.dumps({"config": "cfg4: N=16384, 1024 bins, B=1024, one channel per GPU, NCCL max-hold reduce per step",
                          "n_gpus": world, "steps": steps, "calls_per_step": calls, "ms_per_step": ms / steps,
                          "Msamples_per_s": world * steps * samples / (ms * 1e-3) / 1e6,
                          "per_gpu_Msamples_per_s": steps * samples / (ms * 1e-3) / 1e6,
                          "reduced_maxhold_is_elementwise_max": ok}))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
