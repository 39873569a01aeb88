#!/usr/bin/env bash
# GPU session r04i: host-generated sample tables through the concurrent generator -- the tests that compare them with the device generator's frames; e2e of configs[1]
set -u
O=gpurun_out; mkdir -p $O
timeout 100 python -m pytest tests/test_gpu_frame_overlap.py tests/test_gpu_frames_in_flight.py tests/test_gpu_parity.py -q -m gpu -k "tables or flight or pipeline" -p no:cacheprovider > $O/r04i_host_table_tests.log 2>&1; tail -3 $O/r04i_host_table_tests.log
timeout 60 python bench.py --workload c2 --no-cpu-baseline --steps 5 > $O/r04i_bench_c2.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/r04i_bench_c2.json')); print('c2', d['value'], d['e2e']['value'], d['frames_in_flight'])"
