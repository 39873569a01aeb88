set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_frames_in_flight.py -x -q -m gpu > gpurun_out/r04a_fif_tests.log 2>&1; tail -15 gpurun_out/r04a_fif_tests.log
for L in 1 2 3 4; do timeout 300 python scripts/part_probe.py c4 6 parts=1,8 FramesInFlight=$L >> gpurun_out/r04a_part_probe_c4.log 2>&1; done
cat gpurun_out/r04a_part_probe_c4.log
