#!/usr/bin/env bash
# GPU session r04c (--gpus 2): frames in flight across ranks -- single-process communicator test, bench.py configs[3] at N = 2 under torchrun (pipeline of 3 / plain), C++ example with inflight=3
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_frames_in_flight.py -q -m gpu -p no:cacheprovider > $O/r04c_multi_tests.log 2>&1; echo "pytest rc=$?" >> $O/r04c_multi_tests.log; tail -4 $O/r04c_multi_tests.log
run() { n=$1; wl=$2; steps=$3; tag=$4; shift 4
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$n bench.py --gpus $n --workload $wl --steps $steps --warmup 3 "$@" > $O/r04c_bench_${wl}_n$n$tag.json 2> $O/r04c_bench_${wl}_n$n$tag.err
  python - $O/r04c_bench_${wl}_n$n$tag.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(d["config"]["workload"][:40], "N", d["n_gpus"], "fif", d["frames_in_flight"], round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 2), "ms  e2e", round(d["e2e"]["value"], 1), " frac", round(d["roofline"]["frac"], 3), d["clocks"], d["image_mean_srgb8"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run 2 c4 10 ""
run 2 c4 10 _fif1 --frames-in-flight 1
g++ -std=c++17 -O2 examples/multi_gpu.cpp -Iinclude -Lcudatracerlib_b200 -lctl_b200 -Wl,-rpath,$PWD/cudatracerlib_b200 -o examples/ctl_multi_gpu 2> $O/r04c_example_build.err
timeout 300 examples/ctl_multi_gpu c4 gpus=2 frames=10 inflight=3 check > $O/r04c_example_c4_n2.json 2> $O/r04c_example_c4_n2.err; cat $O/r04c_example_c4_n2.json | cut -c1-500; tail -3 $O/r04c_example_c4_n2.err
