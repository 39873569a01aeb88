#!/usr/bin/env bash
# GPU session r03j (--gpus 2): multi-GPU tests of the final build (communicator behind the C ABI, frames with lanes), configs[3] at N = 2
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_frame_overlap.py -q -m gpu -p no:cacheprovider > $O/r03j_multi_tests.log 2>&1; echo "pytest rc=$?" >> $O/r03j_multi_tests.log; tail -4 $O/r03j_multi_tests.log
sed -e 's/r02x/r03j/g' scripts/r02x_session.sh > /tmp/r03j_n.sh; bash /tmp/r03j_n.sh 2
g++ -std=c++17 -O2 examples/multi_gpu.cpp -Iinclude -Lcudatracerlib_b200 -lctl_b200 -Wl,-rpath,$PWD/cudatracerlib_b200 -o examples/ctl_multi_gpu 2> $O/r03j_example_build.err
timeout 300 examples/ctl_multi_gpu c5 gpus=2 frames=2 spp=16 check > $O/r03j_example_c5_n2.json 2> $O/r03j_example_c5_n2.err; cat $O/r03j_example_c5_n2.json | cut -c1-400
