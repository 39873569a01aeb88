#!/usr/bin/env bash
# GPU session r03e: whole GPU suite and the default bench lines (both arms) of the final build
set -u
O=gpurun_out; mkdir -p $O
timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider > $O/r03e_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/r03e_gpu_tests.log; tail -6 $O/r03e_gpu_tests.log | cut -c1-300
( time timeout 900 python bench.py > $O/r03e_bench_default.json 2> $O/r03e_bench_default.err ) 2> $O/r03e_bench_default.time
( time timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $O/r03e_bench_reference.json 2> $O/r03e_bench_reference.err ) 2> $O/r03e_bench_reference.time
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r03e_bench_default.json").read().strip().split("\n")[-1])
print(d["config"]["workload"][:50], round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), "traffic", d["roofline"]["traffic"], d["clocks"], d.get("cpu_baseline"), "launches", d["gpu_launches"])
for k, v in d.get("extra", {}).get("configs", {}).items(): print(k, round(v["value"], 1), round(v["ms_per_step"], 2), "e2e", round(v["e2e"]["value"], 1), round(v["roofline"]["frac"], 3))
r = json.loads(open("gpurun_out/r03e_bench_reference.json").read().strip().split("\n")[-1])
print("reference arm:", round(r["value"], 3), r["unit"], r["cpu_baseline"]["sample"][:120])
PY
cat $O/r03e_bench_default.time $O/r03e_bench_reference.time | grep real
