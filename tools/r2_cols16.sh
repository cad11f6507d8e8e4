#!/bin/bash
# 16-column accumulate tiles as a default candidate for N = 1024: the driver's command, e2e and the parity suite under the knob
run() {
  env "$@" python bench.py --no-cpu --no-extras --e2e-frames 200 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$*', 'value', round(d['value']), 'burst', round(d['burst']['value']), 'one_stream', round(d['one_stream']['value']), 'unfolded', round(d['unfolded']['value']), 'e2e', round(d['e2e']['value']), 'lat', d['e2e'].get('cfg1_burst_latency_us'), d['clocks']['sm_mhz'])"
}
run A=0
run FOSPHOR_B200_ACC_COLS=16
run A=1
run FOSPHOR_B200_ACC_COLS=16
FOSPHOR_B200_ACC_COLS=16 timeout 400 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
