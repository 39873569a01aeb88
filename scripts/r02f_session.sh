#!/usr/bin/env bash
# GPU session r02f: full-resolution parity (fixed test), ray-queue staging through TMA (tests under a timeout, then the A/B on configs 2 / 4 / 3)
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_staged.py -q -m gpu -x -k "ray_queue" -p no:cacheprovider > $O/r02f_raytma_tests.log 2>&1; echo "pytest rc=$?" >> $O/r02f_raytma_tests.log
timeout 900 python -m pytest tests/test_gpu_fullres.py -q -m gpu -s -p no:cacheprovider > $O/r02f_fullres_tests.log 2>&1; echo "pytest rc=$?" >> $O/r02f_fullres_tests.log
if grep -q "pytest rc=0" $O/r02f_raytma_tests.log; then
for wl in c2 c4 c3; do for rt in 0 1; do
  timeout 300 python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu-baseline --no-extra --set StagedRayTMA=$rt 2>/dev/null | tail -1 > $O/r02f_raytma_${wl}_$rt.json
  python - $O/r02f_raytma_${wl}_$rt.json $wl $rt <<'PY' >> $O/r02f_raytma_summary.log
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read())
    print(sys.argv[2], "StagedRayTMA", sys.argv[3], round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 2), "ms", d["roofline"]["stage_ms_last_batch"])
except Exception as e:
    print(sys.argv[2], "StagedRayTMA", sys.argv[3], "FAILED", e)
PY
done; done
fi
tail -3 $O/r02f_raytma_tests.log; grep -E "passed|failed|frac " $O/r02f_fullres_tests.log; cat $O/r02f_raytma_summary.log
