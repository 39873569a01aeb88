#!/usr/bin/env bash
# GPU session r02b: first device run of the staged traversal kernel (tests, then the launch-shape A/B on configs 2 and 4)
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_staged.py -q -m gpu -x -p no:cacheprovider > $O/r02b_staged_tests.log 2>&1; echo "pytest rc=$?" >> $O/r02b_staged_tests.log
timeout 900 python scripts/r02_staged_ab.py c2 c4 c4:1024 > $O/r02b_staged_ab.log 2> $O/r02b_staged_ab.err
tail -5 $O/r02b_staged_tests.log; cat $O/r02b_staged_ab.log
