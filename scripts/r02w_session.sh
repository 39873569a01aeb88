#!/usr/bin/env bash
# GPU session r02w (--gpus 8): the north star's scaling target -- configs[3] (1 M triangles) at N = 8 / 4 under torchrun, configs[4] at N = 8, the C++ example on 8 GPUs
set -u
O=gpurun_out; mkdir -p $O
run() { n=$1; wl=$2; steps=$3; shift 3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$n bench.py --gpus $n --workload $wl --steps $steps --warmup 3 "$@" > $O/r02w_bench_${wl}_n$n.json 2> $O/r02w_bench_${wl}_n$n.err
  python - $O/r02w_bench_${wl}_n$n.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(d["config"]["workload"][:40], "N", d["n_gpus"], round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 2), "ms  e2e", round(d["e2e"]["value"], 1), " frac", round(d["roofline"]["frac"], 3), d["clocks"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run 8 c4 20

run 8 c5 3
g++ -std=c++17 -O2 examples/multi_gpu.cpp -Iinclude -Lcudatracerlib_b200 -lctl_b200 -Wl,-rpath,$PWD/cudatracerlib_b200 -o examples/ctl_multi_gpu 2> $O/r02w_example_build.err
timeout 300 examples/ctl_multi_gpu c4 gpus=8 frames=10 check > $O/r02w_example_c4_n8.json 2> $O/r02w_example_c4_n8.err; cat $O/r02w_example_c4_n8.json
