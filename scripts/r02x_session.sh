#!/usr/bin/env bash
# GPU session r02x (--gpus N, N = 2 or 4): configs[3] at N ranks under torchrun
set -u
N=${1:-2}
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N bench.py --gpus $N --workload c4 --steps 15 --warmup 3 > $O/r02x_bench_c4_n$N.json 2> $O/r02x_bench_c4_n$N.err
python - $O/r02x_bench_c4_n$N.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
print(d["config"]["workload"][:40], "N", d["n_gpus"], round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 2), "ms  e2e", round(d["e2e"]["value"], 1), " frac", round(d["roofline"]["frac"], 3), d["clocks"])
PY
