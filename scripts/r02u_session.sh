#!/usr/bin/env bash
# GPU session r02u: compute-sanitizer memcheck + racecheck over the round-2 kernels, builder timing after the allocation fix
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python scripts/sanitize_target.py > $O/r02u_sanitize_plain.log 2>&1; tail -3 $O/r02u_sanitize_plain.log
timeout 1500 compute-sanitizer --tool memcheck python scripts/sanitize_target.py > $O/r02u_sanitizer_memcheck.log 2>&1; tail -4 $O/r02u_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck python scripts/sanitize_target.py > $O/r02u_sanitizer_racecheck.log 2>&1; tail -4 $O/r02u_sanitizer_racecheck.log
CTL_GPU_BUILDER_VERBOSE=1 timeout 1200 python scripts/bvh_build_bench.py > $O/r02u_bvh_build_bench.log 2>&1; grep -v "ctl gpu builder" $O/r02u_bvh_build_bench.log | cut -c1-300
