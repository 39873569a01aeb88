#!/usr/bin/env bash
# First GPU call of round 2 (one GPU, ~12 min): everything that was written after round 1's GPU budget ran out gets its first device run, with the numbers
# the round-2 plan needs.  Usage:
#     /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/r02_first_gpu_session.sh'
# Everything lands in gpurun_out/r02a_*; nothing here is a bench value taken under a profiler.
set -u
O=gpurun_out; mkdir -p $O
# 1. the whole GPU suite (the three files that have never run on a device sort last)
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider > $O/r02a_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/r02a_gpu_tests.log
timeout 300 python -m pytest tests/test_gpu_zy_rebraid.py tests/test_gpu_zz_nlm.py -q -m gpu -p no:cacheprovider > $O/r02a_new_tests.log 2>&1; echo "pytest rc=$?" >> $O/r02a_new_tests.log
# 2. re-braiding A/B on config 4 and 5 (scene-level leaves budget; 0 = off)
for wl in c4 c5; do
  for n in 0 64 256 1024 4096; do
    if [ "$wl" = c5 ] && [ $n != 0 ] && [ $n != 1024 ]; then continue; fi
    CTL_REBRAID=$n timeout 400 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/r02a_rebraid_${wl}_$n.json
    python - $O/r02a_rebraid_${wl}_$n.json $wl $n <<'PY' >> $O/r02a_rebraid_summary.log
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read())
    print(sys.argv[2], "rebraid", sys.argv[3], round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 2), "ms  roofline", round(d["roofline"]["frac"], 3), round(d["roofline"].get("bytes_per_ray", 0)), "B/ray")
except Exception as e:
    print(sys.argv[2], "rebraid", sys.argv[3], "FAILED", e)
PY
  done
done
# 2b. post-build optimisation of the mesh trees (re-insertion + rotations, on by default): A/B against the builder's raw trees
for wl in c2 c4; do
  for rot in 0 8; do
    CTL_SBVH_ROTATE=$rot timeout 400 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/r02a_treeopt_${wl}_$rot.json
    python - $O/r02a_treeopt_${wl}_$rot.json $wl $rot <<'PY' >> $O/r02a_treeopt_summary.log
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read())
    print(sys.argv[2], "CTL_SBVH_ROTATE", sys.argv[3], round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 2), "ms  roofline", round(d["roofline"]["frac"], 3), round(d["roofline"].get("bytes_per_ray", 0)), "B/ray")
except Exception as e:
    print(sys.argv[2], "CTL_SBVH_ROTATE", sys.argv[3], "FAILED", e)
PY
  done
done
# 3. ray-level micro-benchmark
for wl in c2 c4; do timeout 300 python scripts/ray_microbench.py $wl > $O/r02a_ray_microbench_$wl.json 2> $O/r02a_ray_microbench_$wl.err; done
# 4. NonLocalMeansFilter timing at 1920x1080 (CUDA events around the pipeline call) + launch list
timeout 300 python - > $O/r02a_nlm_timing.log 2>&1 <<'PY'
import numpy as np, torch, time
import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import ImagePipeline
w, h = 1920, 1080
s = ctl.Scene("c2", w, h); t = ctl.PathTracer(w, h); t.InitializeScene(s); t.setParameter("PixelVarianceBuffer", 1)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); t.setStream(st.cuda_stream)
for p in range(4): t.DoPass(p == 0)
out = torch.empty(w * h * 4, dtype=torch.uint8, device="cuda")
for name, P in (("nlm, weights recomputed", ImagePipeline(5, 1.0, 0.0, 0.45, 1.0)), ("nlm, stored weights", ImagePipeline(5, 1000.0, 0.0, 0.45, 1.0)), ("box filter", ImagePipeline(0, 0.5, 0.5)), ("none", ImagePipeline(-1))):
    for _ in range(2): ctl.api._check(ctl.api.lib().ctl_apply_image_pipeline(t._ctx, 0.0, P, out.data_ptr(), None, None))
    if "stored" in name: t.DoPass(False)   # pass count advances by one -> the stored weights are applied
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); ctl.api._check(ctl.api.lib().ctl_apply_image_pipeline(t._ctx, 0.0, P, out.data_ptr(), None, None)); e1.record(st); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1):.3f} ms at {w}x{h}")
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_nlm -c 20 --csv --log-file $O/r02a_nlm_launches.csv python - > /dev/null 2>&1 <<'PY'
import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import ImagePipeline
w, h = 1920, 1080
s = ctl.Scene("cornell", w, h); t = ctl.PathTracer(w, h); t.InitializeScene(s); t.setParameter("PixelVarianceBuffer", 1)
for p in range(3): t.DoPass(p == 0)
t.applyImagePipeline(ImagePipeline(5, 25.0, 0.0, 0.45, 1.0))
PY
# 5. the default bench line, last (same command the driver runs)
timeout 600 python bench.py > $O/r02a_bench_default_c2.json 2> $O/r02a_bench_default_c2.err
ls -la $O | tail -30
