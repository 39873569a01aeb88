#!/usr/bin/env bash
# GPU session r02o: drain-mode prefetch of the staged kernel (A/B on the one-GPU stand-in for rank 0 of N), claim size 32 as the default
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_staged.py -q -m gpu -x -p no:cacheprovider > $O/r02o_staged_tests.log 2>&1; tail -3 $O/r02o_staged_tests.log
for v in "OverlapWavefronts=0 TravDrainPrefetch=0" "OverlapWavefronts=0 TravDrainPrefetch=1" "OverlapLanes=2 StagedThreads=128 TravDrainPrefetch=1"; do
  timeout 400 python scripts/part_probe.py c4 5 $v >> $O/r02o_part_probe_c4.log 2>&1
done
for v in "OverlapWavefronts=0 TravDrainPrefetch=0" "OverlapWavefronts=0 TravDrainPrefetch=1"; do timeout 300 python scripts/part_probe.py c2 5 $v >> $O/r02o_part_probe_c4.log 2>&1; done
python - <<'PY'
import json
for l in open('gpurun_out/r02o_part_probe_c4.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d["workload"], d["params"], d["n_parts"], d["ms_part0"], d["efficiency"], d.get("efficiency_max_part"))
PY
