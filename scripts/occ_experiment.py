import os, sys
sys.path.insert(0, "/root/repo")
from cudatracerlib_b200 import Scene, PathTracer
s = Scene("c2", 1920, 1080)
t = PathTracer(1920, 1080); t.InitializeScene(s); t.setParameter("MaxPathLength", 8); t.setParameter("StageTimers", 1); t.setParameter("TraversalKernel", 1)
for bps in (8, 10, 12, 16):
    t.setParameter("TraversalBlocksPerSM", bps)
    best = None
    for i in range(4):
        t.DoPass(i == 0); t.synchronize(); ms, _ = t.stageTimes()
        if best is None or ms[1] + ms[3] < best[0]: best = (ms[1] + ms[3], ms[1], ms[3])
    print(os.environ.get("CTL_B200_LIB", "default"), "blocks/SM", bps, best)
