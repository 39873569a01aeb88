"""How much would ray sorting buy? Capture the extension rays of one bounce, reorder them on the host, time ctl_intersect."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cudatracerlib_b200 import Scene, PathTracer, RAY_DTYPE
from bench import WORKLOADS

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
kind, w, h, spp, depth, _ = WORKLOADS[wl]
s = Scene(kind, w, h)
t = PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", depth)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); t.setStream(st.cuda_stream)

def part1by2(x):
    x = x.astype(np.uint64) & 0x3ff
    x = (x | (x << 16)) & 0x30000ff; x = (x | (x << 8)) & 0x300f00f; x = (x | (x << 4)) & 0x30c30c3; x = (x | (x << 2)) & 0x9249249
    return x

def time_rays(rays, any_hit, reps=5):
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 8).copy()).cuda()
    d_res = torch.zeros(len(rays), 4, dtype=torch.int32, device="cuda")
    for _ in range(2): t.intersect_device(len(rays), d_rays.data_ptr(), d_res.data_ptr(), any_hit, st.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): t.intersect_device(len(rays), d_rays.data_ptr(), d_res.data_ptr(), any_hit, st.cuda_stream)
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for kern in (1, 0):
    t.setParameter("TraversalKernel", kern)
    for bounce in (1, 3):
        t.setParameter("CaptureBounce", bounce)
        t.DoPass(True); t.synchronize()
        rays = t.capturedRays(w * h)
        t.setParameter("CaptureBounce", 0)
        lo = np.array(list(s.view.box_min)); hi = np.array(list(s.view.box_max))
        q = np.clip(((rays["o"] - lo) / (hi - lo) * 1024).astype(np.int64), 0, 1023)
        morton = part1by2(q[:, 0]) | (part1by2(q[:, 1]) << 1) | (part1by2(q[:, 2]) << 2)
        octant = ((rays["d"][:, 0] < 0).astype(np.uint64) | ((rays["d"][:, 1] < 0).astype(np.uint64) << 1) | ((rays["d"][:, 2] < 0).astype(np.uint64) << 2))
        base = time_rays(rays, False)
        out = [f"kernel {kern} bounce {bounce}: n={len(rays)} unsorted {base:.3f} ms"]
        for name, key in (("morton30", morton), ("morton15", morton >> 15), ("morton30+oct", (morton << 3) | octant), ("oct+morton30", (octant << 30) | morton), ("morton15+oct", ((morton >> 15) << 3) | octant),
                          ("morton9+oct", ((morton >> 21) << 3) | octant), ("random", np.random.default_rng(1).permutation(len(rays)).astype(np.uint64))):
            perm = np.argsort(key, kind="stable")
            out.append(f"{name} {time_rays(rays[perm], False):.3f}")
        print("  ".join(out), flush=True)
