#!/usr/bin/env bash
# GPU session r03m: straggler deferral inside a wavefront -- tests first (short timeouts), then the one-GPU stand-in for rank 0 of N
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_defer.py -q -m gpu -x -p no:cacheprovider -k "cornell7 or oracle" > $O/r03m_defer_tests_small.log 2>&1; echo "rc=$?" >> $O/r03m_defer_tests_small.log; tail -6 $O/r03m_defer_tests_small.log | cut -c1-300
if grep -q "rc=0" $O/r03m_defer_tests_small.log; then
  timeout 900 python -m pytest tests/test_gpu_defer.py -q -m gpu -x -p no:cacheprovider > $O/r03m_defer_tests.log 2>&1; echo "rc=$?" >> $O/r03m_defer_tests.log; tail -6 $O/r03m_defer_tests.log | cut -c1-300
  if grep -q "rc=0" $O/r03m_defer_tests.log; then
    for v in "DeferStragglers=0" "DeferStragglers=1" "DeferStragglers=1 HandOverDrain=8" "DeferStragglers=1 HandOverDrain=32" "DeferStragglers=1 DeferMaxLag=1" "DeferStragglers=1 DeferMaxLag=2"; do timeout 400 python scripts/part_probe.py c4 5 $v >> $O/r03m_part_probe_c4.log 2>&1; done
    python - <<'PY'
import json
for l in open('gpurun_out/r03m_part_probe_c4.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d["workload"], d["params"], d["n_parts"], d["ms_part0"], d["efficiency"], d.get("efficiency_max_part"))
PY
  fi
fi
