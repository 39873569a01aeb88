#!/usr/bin/env bash
# GPU session r02v: compute-sanitizer memcheck + racecheck over the round-2 kernels
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python scripts/sanitize_target.py > $O/r02v_sanitize_plain.log 2>&1; tail -3 $O/r02v_sanitize_plain.log
timeout 1500 compute-sanitizer --tool memcheck python scripts/sanitize_target.py > $O/r02v_sanitizer_memcheck.log 2>&1; tail -4 $O/r02v_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck python scripts/sanitize_target.py > $O/r02v_sanitizer_racecheck.log 2>&1; tail -4 $O/r02v_sanitizer_racecheck.log
