#!/usr/bin/env bash
# GPU session r04f (--gpus 8): configs[3] at N = 8 after the e2e warm-up fix (every slot's pinned table sets allocated before the timed region)
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r04f_bench_c4_n8.json 2> $O/r04f_bench_c4_n8.err
python - $O/r04f_bench_c4_n8.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
print("N", d["n_gpus"], "fif", d["frames_in_flight"], round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 2), "ms  e2e", round(d["e2e"]["value"], 1), d["e2e"]["steps"], " frac", round(d["roofline"]["frac"], 3), d["clocks"], d["wall_s_timed_region"])
PY
