#!/bin/bash
# one GPU call: the driver's bench command (both arms) and the ncu launch list of the same command (short)
mkdir -p gpurun_out
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fft_power|accumulate|count_|update_kernel|fill_kernel|export_maxhold|publish_rows" -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --passes 2 --e2e-frames 8 --no-extras --no-cpu > gpurun_out/r2_ncu_bench.log 2>&1
