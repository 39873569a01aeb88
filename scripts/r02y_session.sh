#!/usr/bin/env bash
# GPU session r02y: KEY_Regularization on the device
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_regularization.py -q -m gpu -x -s -p no:cacheprovider > $O/r02y_regularization_tests.log 2>&1; echo "pytest rc=$?" >> $O/r02y_regularization_tests.log; grep -E "frac vs|passed|failed|Error|assert" $O/r02y_regularization_tests.log | tail -20 | cut -c1-300
