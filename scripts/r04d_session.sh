#!/usr/bin/env bash
# GPU session r04d (--gpus 8): configs[3] at N = 8 under torchrun with the steps as a pipeline of 3 frames in flight; C++ example with inflight=3 on 8 GPUs
set -u
O=gpurun_out; mkdir -p $O
run() { n=$1; wl=$2; steps=$3; tag=$4; shift 4
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$n bench.py --gpus $n --workload $wl --steps $steps --warmup 3 "$@" > $O/r04d_bench_${wl}_n$n$tag.json 2> $O/r04d_bench_${wl}_n$n$tag.err
  python - $O/r04d_bench_${wl}_n$n$tag.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(d["config"]["workload"][:40], "N", d["n_gpus"], "fif", d["frames_in_flight"], round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 2), "ms  e2e", round(d["e2e"]["value"], 1), " frac", round(d["roofline"]["frac"], 3), d["clocks"], d["image_mean_srgb8"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run 8 c4 20 ""
g++ -std=c++17 -O2 examples/multi_gpu.cpp -Iinclude -Lcudatracerlib_b200 -lctl_b200 -Wl,-rpath,$PWD/cudatracerlib_b200 -o examples/ctl_multi_gpu 2> $O/r04d_example_build.err
timeout 300 examples/ctl_multi_gpu c4 gpus=8 frames=20 inflight=3 check > $O/r04d_example_c4_n8.json 2> $O/r04d_example_c4_n8.err; cat $O/r04d_example_c4_n8.json | cut -c1-500; tail -3 $O/r04d_example_c4_n8.err
