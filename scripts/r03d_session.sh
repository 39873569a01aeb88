#!/usr/bin/env bash
# GPU session r03d: fewer rays in flight = shorter ray latency = shorter drain?  resident threads per SM of the traversal launches, plain frames
set -u
O=gpurun_out; mkdir -p $O
for v in "StagedResidentThreads=384" "StagedResidentThreads=512" "StagedResidentThreads=640" "StagedResidentThreads=768" "StagedResidentThreads=896" "HandOver=1 StagedResidentThreads=512" "HandOver=1 StagedResidentThreads=768"; do timeout 400 python scripts/part_probe.py c4 5 $v >> $O/r03d_part_probe_c4.log 2>&1; done
python - <<'PY'
import json
for l in open('gpurun_out/r03d_part_probe_c4.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d["workload"], d["params"], d["n_parts"], d["ms_part0"], d["efficiency"])
PY
