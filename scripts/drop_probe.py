"""Experiment (build variants libctl_drop*.so only -- their images are WRONG): how long a frame of part 0 of 8 takes when a draining launch abandons its
last rays after a few iterations, and how many rays that touches.  Upper bound of what deferring stragglers could gain."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cudatracerlib_b200 import Scene, PathTracer, TILE, lib
from bench import WORKLOADS
wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
kind, w, h, spp, depth, _ = WORKLOADS[wl]
scene = Scene(kind, w, h)
t = PathTracer(w, h); t.InitializeScene(scene); t.setParameter("MaxPathLength", depth)
stream = torch.cuda.Stream(); t.setStream(stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for n_parts in (1, 8):
    def frame(): t.DoFrame(spp, 8, tile=(TILE, TILE), part=0, n_parts=n_parts)
    for _ in range(2): frame()
    torch.cuda.synchronize(); ms = []
    for _ in range(5):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); frame(); b.record(stream); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    rec = {"lib": os.path.basename(os.environ.get("CTL_B200_LIB", "default")), "workload": wl, "n_parts": n_parts, "ms": round(sorted(ms)[2], 3), "rays": t.getRaysInLastPass()}
    if hasattr(lib(), "ctl_debug_counters"):
        ctr = np.zeros(4 * 257, np.uint32); lib().ctl_debug_counters.argtypes = [C.c_void_p, C.c_void_p, C.c_int]; lib().ctl_debug_counters(t._ctx, ctr.ctypes.data, len(ctr))
        rec["dropped_per_launch"] = [int(ctr[2 * 257 + 2 * b + 1]) for b in range(depth)]; rec["ext_queue"] = [int(ctr[b]) for b in range(depth)]
    print(json.dumps(rec), flush=True)
t.close()
