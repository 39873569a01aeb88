#!/usr/bin/env bash
# GPU session r04e: frames in flight -- resident threads / block size / depth of the pipeline on the one-GPU stand-in; e2e warm-up fix checked on configs[1]
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python scripts/fif_probe.py c4 12 FramesInFlight=3 FramesInFlight=3,StagedResidentThreads=768 FramesInFlight=3,StagedResidentThreads=512 FramesInFlight=2,StagedResidentThreads=512 FramesInFlight=6 FramesInFlight=3,StagedThreads=64 FramesInFlight=3,TravChunk=64 > $O/r04e_fif_probe_c4.log 2>&1; cat $O/r04e_fif_probe_c4.log
timeout 300 python bench.py --workload c2 --no-cpu-baseline --steps 20 > $O/r04e_bench_c2.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/r04e_bench_c2.json')); print('c2', d['value'], d['e2e']['value'], d['frames_in_flight'])"
