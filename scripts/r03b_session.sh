#!/usr/bin/env bash
# GPU session r03b: upper bound of straggler deferral (experiment variants that ABANDON a draining launch's last rays; images wrong by design)
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python scripts/drop_probe.py c4 >> $O/r03b_drop_probe.log 2>&1
for v in i4_l4 i4_l8 i16_l4 i16_l32; do CTL_B200_LIB=$PWD/build_variants/libctl_drop_$v.so timeout 300 python scripts/drop_probe.py c4 >> $O/r03b_drop_probe.log 2>&1; done
cut -c1-400 $O/r03b_drop_probe.log
