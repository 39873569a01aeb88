#!/usr/bin/env bash
# GPU session r03l: configs[4] at full resolution on lanes against the oracle
set -u
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_fullres.py -q -m gpu -x -s -p no:cacheprovider -k "config5" > $O/r03l_c5_fullres.log 2>&1; echo "pytest rc=$?" >> $O/r03l_c5_fullres.log; grep -E "^c5|passed|failed|assert|Error" $O/r03l_c5_fullres.log | tail -8 | cut -c1-300
