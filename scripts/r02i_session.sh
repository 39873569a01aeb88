#!/usr/bin/env bash
# GPU session r02i: ncu --set full of the SHIPPING kernels (staged traversal on configs 4 / 2, per-class shade on config 3), the launch list of the
# default bench command, and the one-GPU stand-in for rank 0 of N (scripts/part_probe.py) as the baseline of the scaling work
set -u
O=gpurun_out; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
# second wavefront of two: 9 staged launches per wavefront (bounce 0, 7 fused, last shadow) -> skip 10 = bounces 1 and 2 of wavefront 2
timeout 900 $NCU -k regex:k_intersect_staged -s 10 -c 2 -f -o $O/prof_r02i_c4_trav python scripts/profile_target.py c4 2 8 > $O/r02i_ncu_c4.log 2>&1; echo "c4 rc=$?"
timeout 600 $NCU -k regex:k_intersect_staged -s 10 -c 2 -f -o $O/prof_r02i_c2_trav python scripts/profile_target.py c2 2 8 > $O/r02i_ncu_c2.log 2>&1; echo "c2 rc=$?"
# shade: every per-class launch of bounces 0 and 1 of the second wavefront
timeout 600 $NCU -k regex:k_shade -s 32 -c 8 -f -o $O/prof_r02i_c3_shade python scripts/profile_target.py c3 2 8 > $O/r02i_ncu_c3.log 2>&1; echo "c3 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02i_launches_bench_py_c4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > $O/r02i_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 600 python scripts/part_probe.py c4 5 > $O/r02i_part_probe_c4.log 2>&1; cat $O/r02i_part_probe_c4.log
timeout 300 python scripts/part_probe.py c4 5 batch=4 > $O/r02i_part_probe_c4_b4.log 2>&1; cat $O/r02i_part_probe_c4_b4.log
ls -la $O | tail -12
