#!/usr/bin/env bash
# GPU session r02g (--gpus 2): communicator behind the C ABI -- multi-GPU tests, C++ example, bench.py under torchrun at N = 2 and N = 1 on the same box
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -x -p no:cacheprovider > $O/r02g_multi_tests.log 2>&1; echo "pytest rc=$?" >> $O/r02g_multi_tests.log
g++ -std=c++17 -O2 examples/multi_gpu.cpp -Iinclude -Lcudatracerlib_b200 -lctl_b200 -Wl,-rpath,$PWD/cudatracerlib_b200 -o examples/ctl_multi_gpu 2> $O/r02g_example_build.err
timeout 300 examples/ctl_multi_gpu c4 gpus=2 frames=5 check > $O/r02g_example_c4_n2.json 2> $O/r02g_example_c4_n2.err
timeout 300 examples/ctl_multi_gpu c4 gpus=1 frames=5 > $O/r02g_example_c4_n1.json 2>> $O/r02g_example_c4_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02g_bench_c4_n2.json 2> $O/r02g_bench_c4_n2.err
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-extra --no-cpu-baseline > $O/r02g_bench_c4_n1.json 2> $O/r02g_bench_c4_n1.err
tail -3 $O/r02g_multi_tests.log; cat $O/r02g_example_c4_n2.json $O/r02g_example_c4_n1.json; for f in $O/r02g_bench_c4_n2.json $O/r02g_bench_c4_n1.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().split('\n')[-1]); print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), d['clocks'])"; done
