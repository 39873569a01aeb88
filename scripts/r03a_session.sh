#!/usr/bin/env bash
# GPU session r03a: numbers of the shipping build for the secondary consumers -- WavefrontPathTracer drop-in (configs 2 / 4), ray-level micro-benchmark
set -u
O=gpurun_out; mkdir -p $O
for wl in c2 c4; do
  timeout 600 python scripts/wpt_bench.py $wl > $O/r03a_wavefront_pt_bench_$wl.json 2> $O/r03a_wavefront_pt_bench_$wl.err; tail -c 900 $O/r03a_wavefront_pt_bench_$wl.json; echo
  timeout 600 python scripts/ray_microbench.py $wl > $O/r03a_ray_microbench_$wl.json 2> $O/r03a_ray_microbench_$wl.err; cut -c1-330 $O/r03a_ray_microbench_$wl.json
done
