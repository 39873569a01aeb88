#!/usr/bin/env bash
# GPU session r02m: is the 1/8-image slowdown the claim size (32 rays per atomic on small queues vs 128 on large ones)?  + GPU builder re-test (auto mode, stats)
set -u
O=gpurun_out; mkdir -p $O
for v in "OverlapWavefronts=0 TravChunk=32" "OverlapWavefronts=0 TravChunk=64" "OverlapWavefronts=0 TravChunk=128" "OverlapWavefronts=0 TravChunk=256" "OverlapLanes=2 StagedThreads=128 TravChunk=128" "OverlapLanes=2 StagedThreads=128 TravChunk=64"; do
  timeout 400 python scripts/part_probe.py c4 5 $v >> $O/r02m_part_probe_c4.log 2>&1
done
python - <<'PY'
import json
for l in open('gpurun_out/r02m_part_probe_c4.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d["params"], d["n_parts"], d["ms_part0"], d["efficiency"], d.get("efficiency_max_part"))
PY
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "bvh_build" > $O/r02m_bvh_tests.log 2>&1; echo "pytest rc=$?" >> $O/r02m_bvh_tests.log; tail -5 $O/r02m_bvh_tests.log
CTL_GPU_BUILDER_VERBOSE=1 timeout 900 python scripts/bvh_build_bench.py > $O/r02m_bvh_build_bench.log 2>&1; cat $O/r02m_bvh_build_bench.log | cut -c1-300
