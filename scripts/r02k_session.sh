#!/usr/bin/env bash
# GPU session r02k: the agglomerative GPU builder (tests, quality / build-time A/B against the LBVH and the CPU tree) and the lane / block-size sweep
# of the overlapped frame on the one-GPU stand-in for rank 0 of 8
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "bvh_build" > $O/r02k_bvh_tests.log 2>&1; echo "pytest rc=$?" >> $O/r02k_bvh_tests.log; tail -15 $O/r02k_bvh_tests.log
timeout 900 python scripts/bvh_build_bench.py 8 32 > $O/r02k_bvh_build_bench.log 2>&1; cat $O/r02k_bvh_build_bench.log
timeout 600 python -m pytest tests/test_gpu_frame_overlap.py -q -m gpu -x -p no:cacheprovider > $O/r02k_overlap_tests.log 2>&1; tail -3 $O/r02k_overlap_tests.log
for v in "OverlapLanes=2 StagedThreads=64" "OverlapLanes=2 StagedThreads=32" "OverlapLanes=4 StagedThreads=128" "OverlapLanes=4 StagedThreads=32" "OverlapLanes=3 StagedThreads=64"; do
  timeout 400 python scripts/part_probe.py c4 5 $v >> $O/r02k_part_probe_c4.log 2>&1
done
cat $O/r02k_part_probe_c4.log
