#!/bin/bash
# the driver's command (1.2 s timed, board at its power cap) under a few knobs: does anything lower the energy per sample?
run() {
  env "$@" python bench.py --no-cpu --no-extras --e2e-frames 40 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$*', round(d['value']), 'burst', round(d['burst']['value']), d['clocks']['sm_mhz'], d['clocks']['power_w_median'], d['clocks']['reasons'])"
}
run A=0
run FOSPHOR_B200_FFT_CTAS=2
run FOSPHOR_B200_OVERLAP=0
run FOSPHOR_B200_ACC_COLS=16
run FOSPHOR_B200_L2_HINTS=1
run FOSPHOR_B200_FFT_VARIANT=1
run A=1
