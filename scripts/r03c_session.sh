#!/usr/bin/env bash
# GPU session r03c: ray hand-over between the launches of two half-wavefronts -- tests first (short timeouts: a traversal bug is a hang), then the one-GPU stand-in
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_handover.py -q -m gpu -x -p no:cacheprovider -k "cornell7 or oracle" > $O/r03c_handover_tests_small.log 2>&1; echo "rc=$?" >> $O/r03c_handover_tests_small.log; tail -5 $O/r03c_handover_tests_small.log | cut -c1-300
if grep -q "rc=0" $O/r03c_handover_tests_small.log; then
  timeout 600 python -m pytest tests/test_gpu_handover.py -q -m gpu -x -p no:cacheprovider > $O/r03c_handover_tests.log 2>&1; echo "rc=$?" >> $O/r03c_handover_tests.log; tail -5 $O/r03c_handover_tests.log | cut -c1-300
  if grep -q "rc=0" $O/r03c_handover_tests.log; then
    for v in "HandOver=0" "HandOver=1" "HandOver=1 HandOverDrain=8" "HandOver=1 HandOverDrain=32" "HandOver=1 HandOverDrain=64"; do timeout 400 python scripts/part_probe.py c4 5 $v >> $O/r03c_part_probe_c4.log 2>&1; done
    python - <<'PY'
import json
for l in open('gpurun_out/r03c_part_probe_c4.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d["workload"], d["params"], d["n_parts"], d["ms_part0"], d["efficiency"], d.get("efficiency_max_part"))
PY
  fi
fi
