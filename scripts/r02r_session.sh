#!/usr/bin/env bash
# GPU session r02r: more lanes on configs[4] (8 wavefronts per frame), block size with lanes, and configs[3] / [1] sanity with the new defaults
set -u
O=gpurun_out; mkdir -p $O
for v in "OverlapLanes=4" "OverlapLanes=6" "OverlapLanes=8" "OverlapLanes=8 StagedThreads=128" "OverlapLanes=8 StagedThreads=64" "OverlapLanes=8 StagedThreads=128 batch=4" "OverlapLanes=4 StagedThreads=128 batch=16"; do
  timeout 600 python scripts/part_probe.py c5 2 parts=8 $v >> $O/r02r_part_probe_c5.log 2>&1
done
for v in "OverlapLanes=8" "OverlapLanes=8 StagedThreads=128"; do timeout 600 python scripts/part_probe.py c5 2 parts=1 $v >> $O/r02r_part_probe_c5.log 2>&1; done
for v in "StagedThreads=512" "StagedThreads=128"; do timeout 600 python scripts/part_probe.py c4 5 $v >> $O/r02r_part_probe_c5.log 2>&1; timeout 600 python scripts/part_probe.py c2 5 parts=1 $v >> $O/r02r_part_probe_c5.log 2>&1; timeout 600 python scripts/part_probe.py c3 5 parts=1 $v >> $O/r02r_part_probe_c5.log 2>&1; done
python - <<'PY'
import json
for l in open('gpurun_out/r02r_part_probe_c5.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d["workload"], d["params"], d["n_parts"], d["ms_part0"], d["efficiency"])
PY
