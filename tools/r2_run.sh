#!/bin/bash
# one GPU call: all GPU tests, smoke, host-feed phase probe, the driver's bench command (both arms), ncu launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r2_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1
python tools/e2e_probe.py phases 2>/dev/null > gpurun_out/r2_phases.txt
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --passes 2 --e2e-frames 8 --no-extras --no-cpu > gpurun_out/r2_ncu_bench.log 2>&1
tail -3 gpurun_out/r2_tests.txt; cat gpurun_out/r2_smoke.txt | tail -2; cat gpurun_out/r2_phases.txt
