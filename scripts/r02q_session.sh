#!/usr/bin/env bash
# GPU session r02q: lanes that come for free -- configs[4] (64 spp = 8 wavefronts per frame) on part 0 of 8 and on the whole image, 1 .. 4 lanes
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_frame_overlap.py -q -m gpu -x -p no:cacheprovider > $O/r02q_tests.log 2>&1; tail -3 $O/r02q_tests.log
for v in "OverlapWavefronts=0" "OverlapLanes=2" "OverlapLanes=3" "OverlapLanes=4" "OverlapLanes=2 StagedThreads=128" "OverlapLanes=4 StagedThreads=128"; do
  timeout 600 python scripts/part_probe.py c5 2 parts=8 $v >> $O/r02q_part_probe_c5.log 2>&1
done
for v in "OverlapWavefronts=0" "OverlapLanes=2" "OverlapLanes=4"; do timeout 600 python scripts/part_probe.py c5 2 parts=1 $v >> $O/r02q_part_probe_c5.log 2>&1; done
python - <<'PY'
import json
for l in open('gpurun_out/r02q_part_probe_c5.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d["workload"], d["params"], d["n_parts"], d["ms_part0"])
PY
