"""Frames in flight on one GPU standing in for rank 0 of 8 (and the whole image): ms per frame of a pipelined sequence for several parameter sets, one scene build.
Usage: fif_probe.py workload frames "K=V,K=V" "K=V" ...   (each argument one parameter set; FramesInFlight=L among them)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cudatracerlib_b200 import Scene, PathTracer, TILE
from bench import WORKLOADS

wl, frames = sys.argv[1], int(sys.argv[2])
kind, w, h, spp, depth, _ = WORKLOADS[wl]
scene = Scene(kind, w, h)
stream = torch.cuda.Stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for arg in sys.argv[3:]:
    t = PathTracer(w, h); t.InitializeScene(scene); t.setParameter("MaxPathLength", depth); t.setStream(stream.cuda_stream)
    params = dict(kv.split("=") for kv in arg.split(",") if kv)
    fif = int(params.get("FramesInFlight", 1))
    for k, v in params.items(): t.setParameter(k, int(v))
    rec = {"workload": wl, "params": params}
    for n_parts in (1, 8):
        ms = []
        for rep in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for i in range(frames):
                flush.zero_()
                t.submitFrame(spp, spp, tile=(TILE, TILE), part=0, n_parts=n_parts)
                if i >= fif - 1: t.acquireFrame()
            while t.framesInFlight(): t.acquireFrame()
            b.record(stream); torch.cuda.synchronize()
            if rep: ms.append(a.elapsed_time(b) / frames)
        rec[f"ms_parts{n_parts}"] = round(sorted(ms)[len(ms) // 2], 3)
    rec["efficiency"] = round(rec["ms_parts1"] / 8 / rec["ms_parts8"], 4)
    print(json.dumps(rec), flush=True)
    t.close()
