"""WavefrontPathTracer on N GPUs by pass index (cudatracerlib_b200.DistributedPasses): run under torchrun like bench.py.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/wpt_scale.py [workload] [spp]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import cudatracerlib_b200 as ctl
from bench import WORKLOADS
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
kind, w, h, spp, depth, _ = WORKLOADS[wl]
spp = int(sys.argv[2]) if len(sys.argv) > 2 else spp
world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
scene = ctl.Scene(kind, w, h)
t = ctl.WavefrontPathTracer(w, h, device=local); t.InitializeScene(scene); t.setParameter("MaxPathLength", depth)
t.setParameter("PassStride", world); t.setParameter("PassPhase", rank)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); t.setStream(stream.cuda_stream)
accum = torch.zeros(h * w * 7, dtype=torch.float32, device=dev); t.setAccumDevicePtr(accum.data_ptr())
dp = ctl.DistributedPasses(accum, lambda p, new_trace: t.DoPass(new_trace), lambda: 0)
def sync():
    torch.cuda.synchronize()
    if world > 1: dist.barrier(); torch.cuda.synchronize()
for _ in range(3): dp.frame(spp)
sync()
r0 = t.getTotalRays()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 5
e0.record(stream)
for _ in range(steps): dp.frame(spp)
e1.record(stream)
sync()
ms = e0.elapsed_time(e1) / steps
tt = torch.tensor([ms, float(t.getTotalRays() - r0) / steps], dtype=torch.float64, device=dev)
if world > 1:
    tmax = tt.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX); tsum = tt.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    ms, rays = float(tmax[0]), float(tsum[1])
else:
    rays = float(tt[1])
if rank == 0:
    img = accum.cpu().numpy().reshape(h, w, 7)
    print(json.dumps({"metric": "Mrays/s, WavefrontPathTracer, frame shared by pass index", "workload": wl, "n_gpus": world, "spp": spp, "ms_per_frame": ms, "rays_per_frame": rays,
                      "value": rays / ms / 1e3, "unit": "Mrays/s", "scaling": "strong", "weight_sum_min_max": [float(img[..., 6].min()), float(img[..., 6].max())], "mean_rgb": float(img[..., :3].mean() / spp)}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
t.close()
