#!/usr/bin/env bash
# GPU session r02d: full-resolution parity tests, per-class shading (tests + A/B on configs 2 / 3), the default bench line (configs[3] + extras)
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fullres.py tests/test_gpu_staged.py -q -m gpu -x -s -p no:cacheprovider > $O/r02d_new_tests.log 2>&1; echo "pytest rc=$?" >> $O/r02d_new_tests.log
for wl in c3 c2; do for sm in 0 1; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-extra --set ShadeMode=$sm 2>/dev/null | tail -1 > $O/r02d_shade_${wl}_$sm.json
  python - $O/r02d_shade_${wl}_$sm.json $wl $sm <<'PY' >> $O/r02d_shade_summary.log
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read())
    print(sys.argv[2], "ShadeMode", sys.argv[3], round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 2), "ms", d["roofline"]["stage_ms_last_batch"])
except Exception as e:
    print(sys.argv[2], "ShadeMode", sys.argv[3], "FAILED", e)
PY
done; done
( time timeout 900 python bench.py > $O/r02d_bench_default.json 2> $O/r02d_bench_default.err ) 2> $O/r02d_bench_default.time
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02d_bench_reference.json 2> $O/r02d_bench_reference.err ) 2> $O/r02d_bench_reference.time
tail -5 $O/r02d_new_tests.log; cat $O/r02d_shade_summary.log; cat $O/r02d_bench_default.time $O/r02d_bench_reference.time
