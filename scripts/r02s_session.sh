#!/usr/bin/env bash
# GPU session r02s: frame tests (lanes, host tables), the whole GPU suite, the default bench line with the new defaults (128-thread blocks, lanes, batch 16 on configs[4])
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_frame_overlap.py -q -m gpu -x -p no:cacheprovider > $O/r02s_frame_tests.log 2>&1; tail -3 $O/r02s_frame_tests.log
( time timeout 900 python bench.py > $O/r02s_bench_default.json 2> $O/r02s_bench_default.err ) 2> $O/r02s_bench_default.time
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02s_bench_default.json").read().strip().split("\n")[-1])
print(d["config"]["workload"][:50], round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), d["clocks"], d.get("cpu_baseline"))
for k, v in d.get("extra", {}).get("configs", {}).items(): print(k, round(v["value"], 1), round(v["ms_per_step"], 2), "e2e", round(v["e2e"]["value"], 1), round(v["roofline"]["frac"], 3), v.get("wavefront_lanes"), v.get("passes_per_wavefront"))
PY
cat $O/r02s_bench_default.time; tail -3 $O/r02s_bench_default.err
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > $O/r02s_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/r02s_gpu_tests.log; tail -6 $O/r02s_gpu_tests.log
