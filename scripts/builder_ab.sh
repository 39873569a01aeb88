#!/usr/bin/env bash
# A/B of mesh-BVH builder settings on the GPU: bench.py Mrays/s for each (builder, alpha, ct) on the given workloads
for w in "$@"; do
  for cfg in "sah 1e-5 1" "sbvh 1e-5 1" "sbvh 10 1" "sbvh 10 0.5" "sbvh 1e-5 0.5" "sbvh 1e-2 1"; do
    set -- $cfg
    CTL_BVH_BUILDER=$1 CTL_SBVH_ALPHA=$2 CTL_SBVH_CT=$3 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print('$w builder=$1 alpha=$2 ct=$3', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms roofline', round(d['roofline']['frac'],3), round(d['roofline']['bytes_per_ray']), 'B/ray')"
  done
done
