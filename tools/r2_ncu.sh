#!/bin/bash
# ncu --set full captures of the hot kernels (round 2): cfg2 (one stream: kernels alone), cfg4 FFT, N=8192 / 4096 FFT.
# Summaries are made on the box (profiles/summarize_ncu.py); only the cfg2 report travels back (64 MiB limit).
mkdir -p gpurun_out /tmp/ncu
export FOSPHOR_B200_OVERLAP=0
ncu --set full --clock-control none --import-source on -k regex:"fft_power_stream|accumulate_fused" -s 2 -c 2 -f -o gpurun_out/r2_full_cfg2 python tools/run_config.py 1024 256 1024 128 1024 1024 3 > gpurun_out/r2_ncu1.log 2>&1
ncu --set full --clock-control none -k regex:"fft_power|accumulate_fused" -s 2 -c 2 -f -o /tmp/ncu/n16384 python tools/run_config.py 16384 1024 1024 16 1024 16384 2 > gpurun_out/r2_ncu2.log 2>&1
ncu --set full --clock-control none -k regex:"fft_power" -s 1 -c 1 -f -o /tmp/ncu/n8192 python tools/run_config.py 8192 256 1024 32 1024 8192 2 > gpurun_out/r2_ncu3.log 2>&1
ncu --set full --clock-control none -k regex:"fft_power" -s 1 -c 1 -f -o /tmp/ncu/n4096 python tools/run_config.py 4096 256 1024 64 1024 4096 2 > gpurun_out/r2_ncu4.log 2>&1
python profiles/summarize_ncu.py gpurun_out/r2_full_cfg2.ncu-rep gpurun_out/r2_ncu_full_cfg2.json
for n in 16384 8192 4096; do python profiles/summarize_ncu.py /tmp/ncu/n$n.ncu-rep gpurun_out/r2_ncu_full_n$n.json; done
ls -la gpurun_out/r2_*
