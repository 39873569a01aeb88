#!/usr/bin/env bash
# GPU session r02p: GPU builder after the sync reduction (tests, build times, quality), the default bench line with this session's defaults (claim size 32)
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_frame_overlap.py -q -m gpu -x -p no:cacheprovider -k "bvh_build or overlap or frame" > $O/r02p_tests.log 2>&1; echo "pytest rc=$?" >> $O/r02p_tests.log; tail -5 $O/r02p_tests.log
CTL_GPU_BUILDER_VERBOSE=1 timeout 900 python scripts/bvh_build_bench.py > $O/r02p_bvh_build_bench.log 2>&1; cat $O/r02p_bvh_build_bench.log | cut -c1-300
( time timeout 900 python bench.py > $O/r02p_bench_default.json 2> $O/r02p_bench_default.err ) 2> $O/r02p_bench_default.time
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02p_bench_default.json").read().strip().split("\n")[-1])
print(d["config"]["workload"][:50], round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), d["clocks"], d.get("cpu_baseline"))
for k, v in d.get("extra", {}).get("configs", {}).items(): print(k, round(v["value"], 1), round(v["ms_per_step"], 2), round(v["roofline"]["frac"], 3))
PY
cat $O/r02p_bench_default.time
