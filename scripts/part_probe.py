"""One GPU standing in for rank 0 of N: frame time of part 0 of n_parts (interleaved tiles), i.e. what each rank of an N-GPU run spends per
frame before the reduce.  ideal = t(1 part) / n_parts; the ratio is the device-side scaling efficiency the tiling itself allows.
Usage: part_probe.py [workload] [frames] [KEY=INT ...]   (extra tracer parameters, e.g. OverlapWavefronts=0, StagedThreads=128; batch=N;
FramesInFlight=L > 1: the frames as a pipeline, ctl_submit_frame_tiled / ctl_acquire_frame, ms = one event bracket over all frames / frames)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cudatracerlib_b200 import Scene, PathTracer, TILE
from bench import WORKLOADS

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 5
parts_list = (1, 2, 4, 8)
params = [a.split("=") for a in sys.argv[3:]]
kind, w, h, spp, depth, _ = WORKLOADS[wl]
scene = Scene(kind, w, h)
t = PathTracer(w, h); t.InitializeScene(scene); t.setParameter("MaxPathLength", depth)
batch = 8
fif = 1
for k, v in params:
    if k == "batch": batch = int(v)
    elif k == "FramesInFlight": fif = int(v); t.setParameter(k, fif)
    elif k == "parts": parts_list = tuple(int(x) for x in v.split(","))
    elif k == "depth": t.setParameter("MaxPathLength", int(v))
    elif k == "fif": pass
    else: t.setParameter(k, int(v))
stream = torch.cuda.Stream(); t.setStream(stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
base = None
for n_parts in parts_list:
    def frame(part=0):
        t.DoFrame(spp, batch, tile=(TILE, TILE), part=part, n_parts=n_parts)   # == what ctl_comm_render_frame runs on a rank before the reduce
    for _ in range(2): frame()
    torch.cuda.synchronize()
    ms = []
    if fif > 1:   # the pipeline: `frames` frames, `fif` in flight, one bracket
        for rep in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for i in range(frames):
                flush.zero_()
                t.submitFrame(spp, batch, tile=(TILE, TILE), part=0, n_parts=n_parts)
                if i >= fif - 1: t.acquireFrame()
            while t.framesInFlight(): t.acquireFrame()
            b.record(stream); torch.cuda.synchronize()
            ms.append(a.elapsed_time(b) / frames)
    else:
      for _ in range(frames):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); frame(); b.record(stream); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    m = sorted(ms)[len(ms) // 2]
    if base is None: base = m
    rec = {"workload": wl, "params": dict(params), "n_parts": n_parts, "ms_part0": round(m, 3), "ideal_ms": round(base / n_parts, 3), "efficiency": round(base / n_parts / m, 4)}
    if n_parts == 8 and len(parts_list) > 1:   # every part in turn: the slowest one is what an 8-GPU frame waits for
        per = []
        for part in range(8):
            frame(part); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.zero_(); a.record(stream); frame(part); b.record(stream); torch.cuda.synchronize()
            per.append(round(a.elapsed_time(b), 3))
        rec["ms_parts"] = per; rec["efficiency_max_part"] = round(base / 8 / max(per), 4)
    print(json.dumps(rec), flush=True)
t.close()
