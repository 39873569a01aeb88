#!/usr/bin/env bash
# GPU session r03g: re-braid budget on configs[3] with the shipping kernel (the SAH's triangle cost was swept with the oracle's counters on the CPU: 1 is the optimum)
set -u
O=gpurun_out; mkdir -p $O
run() { tag=$1; shift; env "$@" timeout 400 python scripts/part_probe.py c4 5 parts=1 2>&1 | tail -1 | sed "s|^|$tag |" >> $O/r03g_tree_params.log; }
run default A=1
run rebraid2048 CTL_REBRAID=2048
run rebraid4096 CTL_REBRAID=4096
run rebraid512 CTL_REBRAID=512
cut -c1-200 $O/r03g_tree_params.log
