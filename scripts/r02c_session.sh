#!/usr/bin/env bash
# GPU session r02c: whole GPU suite with the staged kernel as the default, scheduler sweep, 48-register (1280 resident threads) build variant
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider > $O/r02c_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/r02c_gpu_tests.log
timeout 600 python scripts/r02_tune_sweep.py c2 c4:1024 > $O/r02c_tune_sweep.log 2> $O/r02c_tune_sweep.err
for th in 128 640; do
  CTL_B200_LIB=$PWD/build_variants/libctl_b200_r48.so SWEEP_RESIDENT=1280 SWEEP_THREADS=$th SWEEP_QUICK=1 timeout 300 python scripts/r02_tune_sweep.py c2 c4:1024 >> $O/r02c_r48_variant.log 2>> $O/r02c_r48_variant.err
done
tail -4 $O/r02c_gpu_tests.log; cat $O/r02c_tune_sweep.log $O/r02c_r48_variant.log
