#!/usr/bin/env bash
# GPU session r02l: overlapped lanes with traversal grids that leave block slots for the other lane (StagedResidentThreads < 1024)
set -u
O=gpurun_out; mkdir -p $O
for v in "OverlapLanes=2 StagedThreads=128 StagedResidentThreads=768" "OverlapLanes=2 StagedThreads=128 StagedResidentThreads=896" "OverlapLanes=2 StagedThreads=64 StagedResidentThreads=768" "OverlapLanes=2 StagedThreads=128 StagedResidentThreads=640" "OverlapLanes=3 StagedThreads=128 StagedResidentThreads=512" "OverlapLanes=4 StagedThreads=128 StagedResidentThreads=512"; do
  timeout 400 python scripts/part_probe.py c4 5 $v >> $O/r02l_part_probe_c4.log 2>&1
done
cat $O/r02l_part_probe_c4.log
