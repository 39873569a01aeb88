#!/usr/bin/env bash
# GPU session r03h: smoke() of the final build, the ncu launch list of the default bench command (shares of the step), ncu --set full of the final traversal kernel
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r03h_smoke.log 2>&1; tail -3 $O/r03h_smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r03h_launches_bench_py_c4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > $O/r03h_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_intersect_staged -s 10 -c 2 -f -o $O/prof_r03h_c4_trav python scripts/profile_target.py c4 2 8 > $O/r03h_ncu_c4.log 2>&1; echo "ncu rc=$?"
