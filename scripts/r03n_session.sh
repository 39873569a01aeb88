#!/usr/bin/env bash
# GPU session r03n: compute-sanitizer memcheck + racecheck over the round-2 kernels incl. hand-over, deferral, wavefront frames, Regularization; then the whole suite and the bench lines
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python scripts/sanitize_target.py > $O/r03n_sanitize_plain.log 2>&1; tail -3 $O/r03n_sanitize_plain.log
timeout 1500 compute-sanitizer --tool memcheck python scripts/sanitize_target.py > $O/r03n_sanitizer_memcheck.log 2>&1; tail -4 $O/r03n_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck python scripts/sanitize_target.py > $O/r03n_sanitizer_racecheck.log 2>&1; tail -4 $O/r03n_sanitizer_racecheck.log
sed -e "s/r03e/r03n/g" scripts/r03e_session.sh > /tmp/r03n_e.sh; bash /tmp/r03n_e.sh
