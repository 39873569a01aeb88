#!/usr/bin/env bash
# GPU session r03f: concurrent per-class shade launches (configs 3 / 5)
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_frame_overlap.py -q -m gpu -x -p no:cacheprovider > $O/r03f_tests.log 2>&1; tail -3 $O/r03f_tests.log
for v in "ShadeConcurrent=0" "ShadeConcurrent=1"; do
  timeout 300 python scripts/part_probe.py c3 7 parts=1 $v >> $O/r03f_shade_concurrent.log 2>&1
  timeout 600 python scripts/part_probe.py c5 2 parts=1 $v >> $O/r03f_shade_concurrent.log 2>&1
  timeout 600 python scripts/part_probe.py c5 2 parts=8 batch=16 $v >> $O/r03f_shade_concurrent.log 2>&1
done
cut -c1-200 $O/r03f_shade_concurrent.log
