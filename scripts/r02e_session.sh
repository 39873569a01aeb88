#!/usr/bin/env bash
# GPU session r02e: the whole GPU suite (incl. full-resolution parity), rough-conductor shade launches at 6 / 5 resident blocks per SM (build variants)
set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -s -p no:cacheprovider > $O/r02e_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/r02e_gpu_tests.log
for v in default shade6 shade5; do
  lib=$PWD/cudatracerlib_b200/libctl_b200.so; [ $v != default ] && lib=$PWD/build_variants/libctl_b200_$v.so
  CTL_B200_LIB=$lib timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 > $O/r02e_shade_occ_$v.json
  python - $O/r02e_shade_occ_$v.json $v <<'PY' >> $O/r02e_shade_occ_summary.log
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read())
    print("c3", sys.argv[2], round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 2), "ms", d["roofline"]["stage_ms_last_batch"])
except Exception as e:
    print("c3", sys.argv[2], "FAILED", e)
PY
done
grep -E "passed|failed|frac " $O/r02e_gpu_tests.log | tail -20; cat $O/r02e_shade_occ_summary.log
