#!/bin/bash
# SASS evidence of the sm_100a-native instructions in libfosphor_b200.so -> profiles/sass_summary.txt
LIB=gr-fosphor_b200/libfosphor_b200.so
OUT=profiles/sass_summary.txt
cuobjdump -sass $LIB > /tmp/fosphor.sass
{
echo "# cuobjdump -sass $LIB  ($(date -u +%Y-%m-%d), nvcc $(nvcc --version | grep -o 'V[0-9][0-9.]*' | tail -1), $(grep -c 'Function :' /tmp/fosphor.sass) kernels, arch $(grep -m1 -o 'sm_[0-9a]*' /tmp/fosphor.sass))"
echo "# whole library: occurrences of the Blackwell / Hopper+ specific mnemonics"
for m in UTMALDG.2D UBLKCP.S.G UBLKPF.L2 SYNCS.ARRIVE SYNCS.PHASECHK SYNCS.EXCH ATOMS.POPC.INC ATOMS FADD2 FFMA2 FMUL2 MUFU.LG2 LDS.64 LDS.128 STS.64 BAR.SYNC "UTC.MMA\|UTCHMMA\|UTCQMMA\|HMMA\|WGMMA"; do
  printf "%-28s %6d\n" "$m" "$(grep -c -- "$m" /tmp/fosphor.sass)"
done
echo
echo "# per hot kernel (instruction lines, then the mnemonics that matter)"
for k in 'fft_power_stream_kernelINS_7FftPlanILi1024ELi32ELi32ELi2EEELb1' 'accumulate_fused_kernelILi8ELi16ELi8ELi256ELi64ELi1ELi1E' 'fft_power_half_stage_kernel' 'fft_power_kernelINS_7FftPlanILi8192ELi8ELi32ELi3EEELb0' 'fft_power_kernelINS_7FftPlanILi4096ELi16ELi16ELi3EEELb0'; do
  f=$(grep -o "Function : [^ ]*$k[^ ]*" /tmp/fosphor.sass | head -1 | sed 's/Function : //')
  [ -z "$f" ] && continue
  cuobjdump -sass -fun "$f" $LIB > /tmp/one.sass 2>/dev/null
  echo "## $f"
  grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" /tmp/one.sass | awk '{print $2}' | sed -E 's/^(UTMALDG|UBLKCP|UBLKPF|SYNCS|ATOMS|MUFU)\.([A-Z0-9]+).*/\1.\2/; t; s/\..*//' | sort | uniq -c | sort -rn | head -14 | awk '{printf "   %6d %s\n", $1, $2}'
done
} > $OUT
cat $OUT | head -60
