"""Small frames-in-flight run for compute-sanitizer memcheck: pipelines of 2 / 3 / 7 frames, device- and host-generated tables, several wavefronts per frame, a part of
several, a resize between two pipelines, then a plain frame on the lanes the pipeline used."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cudatracerlib_b200 import Scene, PathTracer
s = Scene("c3", 128, 72)
t = PathTracer(128, 72); t.InitializeScene(s); t.setParameter("MaxPathLength", 6)
for fif, dev, spp, batch, parts in ((2, 1, 4, 4, 1), (3, 0, 4, 2, 1), (7, 1, 2, 2, 3), (1, 0, 2, 1, 1)):
    t.setParameter("FramesInFlight", fif); t.setParameter("DeviceSampleTables", dev)
    for i in range(2 * fif + 1):
        t.submitFrame(spp, batch, part=parts - 1, n_parts=parts)
        if i >= fif - 1: t.acquireFrame(); t.resolveSRGB8()
    while t.framesInFlight(): t.acquireFrame()
    img = t.readAccumulator()
t.Resize(96, 64); s2 = Scene("soup", 96, 64); t.InitializeScene(s2); t.setParameter("FramesInFlight", 2)
t.submitFrame(2, 2); t.submitFrame(2, 2); t.acquireFrame(); t.acquireFrame()
t.setParameter("OverlapLanes", 4); t.DoFrame(8, 2); t.synchronize()
print("fif ok", float(t.readAccumulator()["rgb"].mean()))
t.close()
