#!/usr/bin/env bash
# GPU session r02j: overlapped wavefront lanes (ctl_render_frame_tiled) -- tests, then the one-GPU stand-in for rank 0 of N with and without the overlap
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_frame_overlap.py tests/test_gpu_multi.py -q -m gpu -x -p no:cacheprovider > $O/r02j_tests.log 2>&1; echo "pytest rc=$?" >> $O/r02j_tests.log; tail -4 $O/r02j_tests.log
for v in "OverlapWavefronts=0" "OverlapWavefronts=1" "OverlapWavefronts=1 StagedThreads=128" "OverlapWavefronts=1 StagedThreads=256" "OverlapWavefronts=0 StagedThreads=128"; do
  timeout 400 python scripts/part_probe.py c4 5 $v >> $O/r02j_part_probe_c4.log 2>&1
done
cat $O/r02j_part_probe_c4.log
for v in "OverlapWavefronts=0" "OverlapWavefronts=1 StagedThreads=128"; do timeout 300 python scripts/part_probe.py c2 5 $v >> $O/r02j_part_probe_c2.log 2>&1; done
cat $O/r02j_part_probe_c2.log
