#!/usr/bin/env bash
# GPU session r02z: marginal cost of one L1 wavefront per lane and node step (build variants with 1 / 2 unused extra loads of the node's line)
set -u
O=gpurun_out; mkdir -p $O
for lib in "" build_variants/libctl_extra1.so build_variants/libctl_extra2.so; do
  for wl in c4 c2; do
    CTL_B200_LIB=${lib:+$PWD/$lib} timeout 400 python scripts/part_probe.py $wl 5 parts=1 2>&1 | tail -1 | sed "s|^|${lib:-default} |" >> $O/r02z_extra_wavefronts.log
  done
done
cat $O/r02z_extra_wavefronts.log | cut -c1-200
CTL_B200_LIB=$PWD/build_variants/libctl_extra1.so timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_intersect_staged -s 10 -c 2 --csv --log-file $O/r02z_ncu_extra1_c4.csv python scripts/profile_target.py c4 2 8 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_intersect_staged -s 10 -c 2 --csv --log-file $O/r02z_ncu_extra0_c4.csv python scripts/profile_target.py c4 2 8 > /dev/null 2>&1
grep -h "k_intersect" $O/r02z_ncu_extra0_c4.csv $O/r02z_ncu_extra1_c4.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-200
