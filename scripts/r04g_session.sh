#!/usr/bin/env bash
# GPU session r04g: final verification of the build -- compute-sanitizer memcheck over the frames-in-flight pipeline, the whole GPU suite, both bench arms, ncu launch list of the bench command
set -u
O=gpurun_out; mkdir -p $O
timeout 240 compute-sanitizer --tool memcheck python scripts/sanitize_fif_target.py > $O/r04g_sanitizer_memcheck_fif.log 2>&1; tail -3 $O/r04g_sanitizer_memcheck_fif.log
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > $O/r04g_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/r04g_gpu_tests.log; tail -6 $O/r04g_gpu_tests.log | cut -c1-300
( time timeout 600 python bench.py > $O/r04g_bench_default.json 2> $O/r04g_bench_default.err ) 2> $O/r04g_bench_default.time
( time timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > $O/r04g_bench_reference.json 2> $O/r04g_bench_reference.err ) 2> $O/r04g_bench_reference.time
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r04g_bench_default.json").read().strip().split("\n")[-1])
print(d["config"]["workload"][:50], round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 1), "fif", d["frames_in_flight"], "frac", round(d["roofline"]["frac"], 3), "traffic", d["roofline"]["traffic"], d["clocks"], d.get("cpu_baseline"), "launches", d["gpu_launches"])
for k, v in d.get("extra", {}).get("configs", {}).items(): print(k, round(v["value"], 1), round(v["ms_per_step"], 2), "e2e", round(v["e2e"]["value"], 1), round(v["roofline"]["frac"], 3))
r = json.loads(open("gpurun_out/r04g_bench_reference.json").read().strip().split("\n")[-1])
print("reference arm:", round(r["value"], 3), r["unit"], r["cpu_baseline"]["sample"][:120])
PY
cat $O/r04g_bench_default.time $O/r04g_bench_reference.time | grep real
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r04g_launches_bench_py_c4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > $O/r04g_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
