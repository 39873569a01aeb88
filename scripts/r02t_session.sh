#!/usr/bin/env bash
# GPU session r02t: triangle pre-splitting in the GPU builders -- tests, quality / build-time table
set -u
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -p no:cacheprovider -k "bvh" > $O/r02t_bvh_tests.log 2>&1; echo "pytest rc=$?" >> $O/r02t_bvh_tests.log; grep -E "pre-split|passed|failed|Error|error" $O/r02t_bvh_tests.log | tail -12 | cut -c1-300
CTL_GPU_BUILDER_VERBOSE=1 timeout 1200 python scripts/bvh_build_bench.py > $O/r02t_bvh_build_bench.log 2>&1; grep -v "ctl gpu builder" $O/r02t_bvh_build_bench.log | cut -c1-300; grep "ctl gpu builder" $O/r02t_bvh_build_bench.log | tail -12 | cut -c1-200
