#!/usr/bin/env bash
# GPU session r03i: experiment -- rays that have been in their lane for many iterations take more node steps per iteration (shorter wall time for the longest rays)
set -u
O=gpurun_out; mkdir -p $O
for lib in a64_b4 a128_b4 a64_b8 a32_b3; do
  CTL_B200_LIB=${lib:+$PWD/build_variants/libctl_boost_$lib.so} timeout 400 python scripts/part_probe.py c4 5 2>&1 | sed "s|^|${lib:-default} |" >> $O/r03i_old_ray_boost.log
done
python - <<'PY'
import json
for l in open('gpurun_out/r03i_old_ray_boost.log'):
    tag, _, js = l.partition(" ")
    try: d=json.loads(js)
    except Exception: print(l.strip()[:200]); continue
    print(tag, d["n_parts"], d["ms_part0"], d["efficiency"])
PY
