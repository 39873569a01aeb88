#!/usr/bin/env bash
# GPU session r03k: WavefrontPathTracer frames on lanes (tests, timing on configs 2 / 4), retry of the ncu launch list of the default bench command
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_wavefront_pt.py -q -m gpu -x -p no:cacheprovider > $O/r03k_wpt_tests.log 2>&1; echo "pytest rc=$?" >> $O/r03k_wpt_tests.log; tail -4 $O/r03k_wpt_tests.log | cut -c1-300
for wl in c2 c4; do
  timeout 600 python scripts/wpt_bench.py $wl > $O/r03k_wavefront_pt_bench_$wl.json 2> $O/r03k_wavefront_pt_bench_$wl.err
  python - $O/r03k_wavefront_pt_bench_$wl.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
print(d["scene"], "passes one by one", round(d["wavefront_pt"]["mrays_s"], 1), "frames on lanes", {k: round(v["mrays_s"], 1) for k, v in d["wavefront_pt_frame"].items()}, "PathTracer", round(d["path_tracer"]["mrays_s"], 1))
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r03k_launches_bench_py_c4.csv python -X faulthandler bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > $O/r03k_bench_under_ncu.log 2>&1; echo "launch list rc=$?"; tail -5 $O/r03k_bench_under_ncu.log | cut -c1-300
